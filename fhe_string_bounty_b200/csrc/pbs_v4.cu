// pbs_v4.cu -- blind rotation, fourth generation: the v3 data path (Fourier key streamed once per SM through a TMA-fed
// shared-memory ring, accumulator master copy in Tensor Memory) with HALF the work per thread: two warps per polynomial,
// 16 FFT points per thread (fft16_core.cuh), so a 4-ciphertext CTA has 16 warps -- 4 per scheduler instead of 2.
//
// Why: ncu on v3 (profiles/r01_pbs_v3_final_b8192_summary.txt) shows FP64 pipe 45 %, LSU 52 %, issue slots 46 %: nothing is
// saturated, the two 255-register warps per scheduler simply cannot cover each other's dependency stalls.  With 16 points
// per thread the kernel fits 128 registers, and the schedulers get twice the warps to pick from, at the price of one more shared-memory exchange per FFT (16 x 16 x 4 instead of 32 x 32).
//
// Same arithmetic definition as pbs_v3.cu (bootstrap.rs:242-364, ggsw.rs:477-598, fft/mod.rs:197-326): the FFT evaluates
// the same 1024 frequencies, only the butterfly order (hence FP64 rounding) differs.
//
// Shared memory per 4-ciphertext CTA: 8 x 17 KiB polynomial tiles (rotated-gather source, then exchange tile, then spectrum
// exchange) + 5 x 16 KiB ring = 216 KiB.  TMEM: 64 columns per ciphertext (accumulator master copy) + 80 columns of
// per-thread FFT twiddles (ncu: the kernel is bound by the LSU pipe, tcgen05.ld is not on it).
// Named barriers: 1..8 one per polynomial (64 threads), 9..12 one per ciphertext (128), 13..15 start-up stagger.
#include "kernels.h"
#include "pbs16_common.cuh"

namespace tb4 {
using namespace tb16k;

constexpr int QPP = 4;                 // frequencies (registers) per ring piece
constexpr int PIECE_CPLX = 2 * QPP * 64;       // one output polynomial's share of a chunk: [sel 2][q QPP][thread 64] = 8 KiB
constexpr int PIECE_BYTES = PIECE_CPLX * 16;
constexpr int CHUNKS_PER_ITER = 16 / QPP;
// ring slots PER OUTPUT POLYNOMIAL: the warps of polynomial w only ever read the c = w half of a chunk, so the key streams through two
// interleaved rings (slot 2 k + w) whose pieces are consumed by 2 * CTS warps each.  1..4 ciphertexts per CTA: 5 + 5 slots = 80 KiB;
// 5 ciphertexts: 3 + 3 slots = 48 KiB next to ten 17 KiB tiles.
constexpr int ring_slots(int cts) { return cts == 5 ? 3 : 5; }

template <int CTS>
struct Smem {
    static constexpr int NS = ring_slots(CTS);
    cplx tile[2 * CTS][kTileCplx];         // 17 KiB per polynomial
    cplx ring[2 * NS][PIECE_CPLX];
    unsigned long long full_bar[2 * NS];
    unsigned int consumed[2 * NS];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem<4>) <= 227 * 1024, "shared memory budget");
static_assert(sizeof(Smem<5>) <= 227 * 1024, "shared memory budget");

// Fourier key, v4 layout: [ggsw i][chunk 16/QPP][out poly c][sel: 0 = row c, 1 = row 1-c][q QPP][thread 64]; register g = QPP*chunk + q
// (QPP = 4 measured 101.9 ms per 8192 against 104.1 ms for QPP = 2; QPP = 8 leaves too few slots for the ciphertexts' stagger)
__device__ __forceinline__ size_t bskf4_index(int i, int chunk, int c, int sel, int q) {
    return ((((size_t)(i * CHUNKS_PER_ITER + chunk) * 2 + c) * 2 + sel) * QPP + q) * 64;
}
// piece k of output polynomial w (k = i * CHUNKS_PER_ITER + chunk): 8 KiB, contiguous in the layout above
__device__ __forceinline__ const cplx *piece_src(const cplx *bskf4, int k, int w) { return bskf4 + ((size_t)k * 2 + w) * PIECE_CPLX; }

template <int CTS>
__global__ void __launch_bounds__(128 * CTS, 1)
pbs_classic_kernel_v4(const uint64_t *__restrict__ lwe_small, const uint32_t *__restrict__ lut_idx, const uint64_t *__restrict__ luts,
                      const cplx *__restrict__ bskf4, const cplx *__restrict__ tbl16, uint64_t *__restrict__ out,
                      const uint32_t *__restrict__ out_slot, int batch, int n, int base_log, int n_iters, int small_is_u16) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TMEM_COLS = CTS >= 4 ? 512 : 256, TW_COL = 64 * CTS, NS = Smem<CTS>::NS, CONSUMERS = 2 * CTS;
    constexpr bool PER_CHUNK = NS <= CHUNKS_PER_ITER;    // the ring cannot hold a whole iteration plus the chunk being waited for
    static_assert(TW_COL + 80 <= TMEM_COLS, "TMEM columns");
    static_assert(3 * CTS <= 15, "named barriers: 1 .. 2 CTS per polynomial, 2 CTS + 1 .. 3 CTS per ciphertext");
    Smem<CTS> &sm = *reinterpret_cast<Smem<CTS> *>(smem_raw);
    const int W = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp -> (ciphertext, polynomial, half): the four warps of a ciphertext sit on the four schedulers (warp id % 4), so every
    // scheduler hosts one warp of each ciphertext; the ciphertexts are started a quarter iteration apart (stagger below)
    const int ctl = W >> 2, w = (W >> 1) & 1, T = ((W & 1) << 5) | lane, P = W >> 1;
    const int ct_raw = blockIdx.x * CTS + ctl;
    const bool live = ct_raw < batch;
    const int ct = live ? ct_raw : batch - 1;        // ragged tail: recompute the last ciphertext, skip the store
    cplx *tile = sm.tile[P];
    const cplx *otile = sm.tile[P ^ 1];
    uint64_t *pb = reinterpret_cast<uint64_t *>(tile);   // the polynomial (2048 words) for the rotated gather
    const PolySync poly_sync{1 + P};
    const int ct_bar = 1 + 2 * CTS + ctl;
    const uint64_t *lwe = lwe_small + (size_t)ct * (n + 1);
    const uint16_t *lwe16 = reinterpret_cast<const uint16_t *>(lwe_small) + (size_t)ct * (n + 1);
    const int total_pieces = n_iters * CHUNKS_PER_ITER;      // per output polynomial

    // ---- one-time setup: twiddle tables, barriers, TMEM, first ring fill ---------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * NS; ++s) { mbar_init(&sm.full_bar[s], 1); sm.consumed[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (W == 0) tmem_alloc<TMEM_COLS>(&sm.tmem_base);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_mine = sm.tmem_base + ((uint32_t)((W & 3) * 32) << 16) + (uint32_t)(ctl * 64);   // lane quarter = warp id % 4
    const TmemTwiddles twd{sm.tmem_base + ((uint32_t)((W & 3) * 32) << 16) + (uint32_t)TW_COL};
    {   // this thread's twiddles -> TMEM (5 x 16 columns: T1[p][T], p = 0..15, then the three pass-2 twiddles)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            uint32_t v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const cplx t = k < 4 ? __ldg(tbl16 + (4 * k + q) * 64 + T) : __ldg(tbl16 + kM + 16 * (q < 3 ? q : 0) + (T & 15));
                v[4 * q] = (uint32_t)__double2loint(t.x); v[4 * q + 1] = (uint32_t)__double2hiint(t.x);
                v[4 * q + 2] = (uint32_t)__double2loint(t.y); v[4 * q + 3] = (uint32_t)__double2hiint(t.y);
            }
            tmem_st16(twd.col + 16 * k, v);
        }
    }
    if (threadIdx.x == 0) {
        const int first = total_pieces < NS ? total_pieces : NS;
        for (int k = 0; k < first; ++k)
            for (int c = 0; c < 2; ++c) {
                mbar_expect_tx(&sm.full_bar[2 * k + c], PIECE_BYTES);
                tma_load_1d(sm.ring[2 * k + c], piece_src(bskf4, k, c), PIECE_BYTES, &sm.full_bar[2 * k + c]);
            }
    }

    // ---- acc <- LUT * X^(-b_hat): registers (own coefficients as u64 bit patterns in re/im), TMEM, shared ------------------------
    double re[16], im[16];
    {
        const uint32_t b_hat = (small_is_u16 ? (uint32_t)__ldg(lwe16 + n) : modulus_switch_2n(__ldg(lwe + n))) & (2 * kN - 1);
        const uint32_t a0 = (2 * kN - b_hat) & (2 * kN - 1);
        const uint64_t *lut = luts + ((size_t)(lut_idx ? lut_idx[ct] : 0) * 2 + w) * kN;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int j = T + 64 * m;
            int s0, s1; bool n0, n1;
            rot_src(j, a0, s0, n0);
            rot_src(j + kM, a0, s1, n1);
            uint64_t v0 = __ldg(lut + s0), v1 = __ldg(lut + s1);
            v0 = n0 ? (uint64_t)0 - v0 : v0;
            v1 = n1 ? (uint64_t)0 - v1 : v1;
            pb[j] = v0; pb[j + kM] = v1;
            re[m] = __longlong_as_double((long long)v0);
            im[m] = __longlong_as_double((long long)v1);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v[16];
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                const unsigned long long a = (unsigned long long)__double_as_longlong(re[4 * k + mm]);
                const unsigned long long b = (unsigned long long)__double_as_longlong(im[4 * k + mm]);
                v[4 * mm] = (uint32_t)a; v[4 * mm + 1] = (uint32_t)(a >> 32);
                v[4 * mm + 2] = (uint32_t)b; v[4 * mm + 3] = (uint32_t)(b >> 32);
            }
            tmem_st16(tmem_mine + 16 * k, v);
        }
        tmem_wait_st();
    }

    // stagger: ciphertext k starts k / CTS of an iteration after ciphertext 0, so that the FP64-heavy and the shared-memory-heavy phases
    // of the ciphertexts sharing a scheduler do not coincide (a clock-based delay measured the same as a barrier hand-shake)
    if (CTS >= 4 && ctl >= 1 && n_iters > 0) {
        const long long t0 = clock64(), delay = (long long)ctl * (20000 / CTS);
        while (clock64() - t0 < delay) { }
    }

    int slot = 0;            // this polynomial's ring position (0 .. NS-1) of this iteration's first piece
    uint32_t phase = 0;

    for (int i = 0; i < n_iters; ++i) {
        const uint32_t a = (small_is_u16 ? (uint32_t)__ldg(lwe16 + i) : modulus_switch_2n(__ldg(lwe + i))) & (2 * kN - 1);   // a == 0 is NOT skipped
        poly_sync();    // the accumulator polynomial is complete in shared memory

        // ct1 = acc * X^a - acc, level-1 signed digit, folded (own coefficients come from the registers)
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int j = T + 64 * m;
            const uint32_t s0 = ((uint32_t)j - a) & (2 * kN - 1);
            const uint32_t s1 = (s0 + kM) & (2 * kN - 1);
            uint64_t r0 = pb[s0 & (kN - 1)], r1 = pb[s1 & (kN - 1)];
            r0 = (s0 >= (uint32_t)kN) ? (uint64_t)0 - r0 : r0;
            r1 = (s1 >= (uint32_t)kN) ? (uint64_t)0 - r1 : r1;
            const uint64_t o0 = (uint64_t)__double_as_longlong(re[m]), o1 = (uint64_t)__double_as_longlong(im[m]);
            re[m] = (double)signed_digit_l1(r0 - o0, base_log);
            im[m] = (double)signed_digit_l1(r1 - o1, base_log);
        }

        fft16_fwd(re, im, tile, twd, T, poly_sync);

        // spectrum exchange between the two polynomials of the ciphertext: park my 16 values in my own exchange-B reader slots
        {
            cplx *wp = tile + xb_rbase(T);
#pragma unroll
            for (int g = 0; g < 16; ++g) { cplx v; v.x = re[g]; v.y = im[g]; wp[xb_roff(g)] = v; }
        }
        bar_sync(ct_bar, 128);

        // out_fft[w] = F_w * G[w][w] + F_{1-w} * G[1-w][w], GGSW pieces from this polynomial's ring.  Nothing is synchronised inside the loop
        // (the compiler is free to run chunk c+1's loads under chunk c's arithmetic); after it lane c counts this warp out of piece c's
        // slot, and the last of the 2 * CTS warps to leave a slot re-arms it with the piece NS chunks ahead (measured against releasing
        // each slot right after its chunk: 104.7 vs 106.0 ms per 8192).
        {
            const cplx *fop = otile + xb_rbase(T);
            int my_slot = 0;
#pragma unroll
            for (int c = 0; c < CHUNKS_PER_ITER; ++c) {
                const int rs = 2 * slot + w;
                if (!mbar_try_wait(&sm.full_bar[rs], phase)) mbar_wait(&sm.full_bar[rs], phase);
                const cplx *pc = sm.ring[rs] + T;
#pragma unroll
                for (int q = 0; q < QPP; ++q) {
                    const int g = QPP * c + q;
                    const cplx A = pc[q * 64], B = pc[(QPP + q) * 64], F = fop[xb_roff(g)];
                    const double fr = re[g], fi = im[g];
                    double orr = DMUL(fr, A.x);
                    orr = DFMA(-fi, A.y, orr);
                    orr = DFMA(F.x, B.x, orr);
                    orr = DFMA(-F.y, B.y, orr);
                    double oi = DMUL(fr, A.y);
                    oi = DFMA(fi, A.x, oi);
                    oi = DFMA(F.x, B.y, oi);
                    oi = DFMA(F.y, B.x, oi);
                    re[g] = orr; im[g] = oi;
                }
                if (PER_CHUNK) {
                    // fewer slots than chunks per iteration (5 ciphertexts per CTA): a slot has to be handed back before the next chunk can
                    // arrive, so this warp counts itself out right away (shared-memory accesses of a warp complete in order: the atomic
                    // follows every lane's loads of this chunk)
                    __syncwarp();
                    if (lane == 0 && atomicAdd(&sm.consumed[rs], 1u) == CONSUMERS - 1) {
                        sm.consumed[rs] = 0;
                        const int k2 = i * CHUNKS_PER_ITER + c + NS;
                        if (k2 < total_pieces) {
                            __threadfence_block();
                            fence_proxy_async();
                            mbar_expect_tx(&sm.full_bar[rs], PIECE_BYTES);
                            tma_load_1d(sm.ring[rs], piece_src(bskf4, k2, w), PIECE_BYTES, &sm.full_bar[rs]);
                        }
                    }
                }
                if (lane == c) my_slot = rs;
                if (++slot == NS) { slot = 0; phase ^= 1u; }
            }
            __syncwarp();     // every lane's loads from the ring have returned (their values fed the arithmetic above)
            if (!PER_CHUNK && lane < CHUNKS_PER_ITER && atomicAdd(&sm.consumed[my_slot], 1u) == CONSUMERS - 1) {
                sm.consumed[my_slot] = 0;
                const int k2 = i * CHUNKS_PER_ITER + lane + NS;
                if (k2 < total_pieces) {
                    __threadfence_block();
                    fence_proxy_async();
                    mbar_expect_tx(&sm.full_bar[my_slot], PIECE_BYTES);
                    tma_load_1d(sm.ring[my_slot], piece_src(bskf4, k2, w), PIECE_BYTES, &sm.full_bar[my_slot]);
                }
            }
        }
        bar_sync(ct_bar, 128);   // the partner polynomial has read my spectrum: the tile is mine again

        fft16_inv(re, im, tile, twd, T, poly_sync);
        poly_sync();    // everyone has read the last exchange: the tile becomes the accumulator polynomial again

        // acc += from_torus(.): master copy in TMEM, new values to registers (next gather's "own") and shared (next rotation)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v[16];
            tmem_ld16(tmem_mine + 16 * k, v);
            tmem_wait_ld();
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                const int m = 4 * k + mm, j = T + 64 * m;
                uint64_t o0 = ((uint64_t)v[4 * mm + 1] << 32) | v[4 * mm];
                uint64_t o1 = ((uint64_t)v[4 * mm + 3] << 32) | v[4 * mm + 2];
                o0 += from_torus_f64(re[m]);
                o1 += from_torus_f64(im[m]);
                v[4 * mm] = (uint32_t)o0; v[4 * mm + 1] = (uint32_t)(o0 >> 32);
                v[4 * mm + 2] = (uint32_t)o1; v[4 * mm + 3] = (uint32_t)(o1 >> 32);
                pb[j] = o0; pb[j + kM] = o1;
                re[m] = __longlong_as_double((long long)o0);
                im[m] = __longlong_as_double((long long)o1);
            }
            tmem_st16(tmem_mine + 16 * k, v);
        }
        tmem_wait_st();
    }

    // sample extraction (coefficient 0) straight from the registers: out[0] = A[0], out[N-j] = -A[j]; body = B[0]
    if (live) {
        uint64_t *o = out + (size_t)(out_slot ? out_slot[ct] : ct) * (kN + 1);
        if (w == 0) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int j = T + 64 * m;
                const uint64_t v0 = (uint64_t)__double_as_longlong(re[m]), v1 = (uint64_t)__double_as_longlong(im[m]);
                if (j == 0) o[0] = v0; else o[kN - j] = (uint64_t)0 - v0;
                o[kN - (j + kM)] = (uint64_t)0 - v1;
            }
        } else if (T == 0) {
            o[kN] = (uint64_t)__double_as_longlong(re[0]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (W == 0) tmem_dealloc<TMEM_COLS>(sm.tmem_base);
}

// std -> Fourier key in the v4 ring layout (64 threads per polynomial; same forward transform as the kernel above)
__global__ void __launch_bounds__(64)
bsk_convert_kernel_v4(const uint64_t *__restrict__ bsk_std, cplx *__restrict__ bskf4, const cplx *__restrict__ tbl16, int n_polys) {
    __shared__ cplx tile[kTileCplx];
    const int qd = blockIdx.x, T = threadIdx.x;
    if (qd >= n_polys) return;
    const int i = qd >> 2, r = (qd >> 1) & 1, c = qd & 1;   // std layout [i][level 1][row r][col c][N]
    const uint64_t *src = bsk_std + (size_t)qd * kN;
    const double scale = 5.293955920339377e-23;              // 2^-74
    double re[16], im[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int j = T + 64 * m;
        re[m] = DMUL((double)(long long)src[j], scale);
        im[m] = DMUL((double)(long long)src[j + kM], scale);
    }
    fft16_fwd(re, im, tile, GlobalTwiddles{tbl16, T}, T, BlockSync{});
    const int sel = (r == c) ? 0 : 1;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        cplx v; v.x = re[g]; v.y = im[g];
        bskf4[bskf4_index(i, g / QPP, c, sel, g % QPP) + T] = v;
    }
}

}  // namespace tb4

namespace tbk {

cudaError_t pbs_v4_configure() {
    cudaError_t e = cudaFuncSetAttribute(tb4::pbs_classic_kernel_v4<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tb4::Smem<5>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tb4::pbs_classic_kernel_v4<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tb4::Smem<4>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tb4::pbs_classic_kernel_v4<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tb4::Smem<2>));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tb4::pbs_classic_kernel_v4<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tb4::Smem<1>));
}

cudaError_t launch_pbs_classic_v4(const uint64_t *lwe_small, const uint32_t *lut_idx, const uint64_t *luts, const void *bskf4,
                                  const void *tbl16, uint64_t *out, const uint32_t *out_slot, int batch, int n, int base_log,
                                  int n_iters, int small_is_u16, int wide_cts, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const tb::cplx *bk = reinterpret_cast<const tb::cplx *>(bskf4), *tb = reinterpret_cast<const tb::cplx *>(tbl16);
    if (batch <= sms)
        tb4::pbs_classic_kernel_v4<1><<<batch, 128, sizeof(tb4::Smem<1>), stream>>>(lwe_small, lut_idx, luts, bk, tb, out, out_slot, batch, n,
                                                                                  base_log, n_iters, small_is_u16);
    else if (batch <= 2 * sms)
        tb4::pbs_classic_kernel_v4<2><<<(batch + 1) / 2, 256, sizeof(tb4::Smem<2>), stream>>>(lwe_small, lut_idx, luts, bk, tb, out, out_slot,
                                                                                            batch, n, base_log, n_iters, small_is_u16);
    else if (wide_cts == 5)
        tb4::pbs_classic_kernel_v4<5><<<(batch + 4) / 5, 640, sizeof(tb4::Smem<5>), stream>>>(lwe_small, lut_idx, luts, bk, tb, out, out_slot,
                                                                                            batch, n, base_log, n_iters, small_is_u16);
    else
        tb4::pbs_classic_kernel_v4<4><<<(batch + 3) / 4, 512, sizeof(tb4::Smem<4>), stream>>>(lwe_small, lut_idx, luts, bk, tb, out, out_slot,
                                                                                            batch, n, base_log, n_iters, small_is_u16);
    return cudaGetLastError();
}

cudaError_t launch_bsk_convert_v4(const uint64_t *bsk_std, void *bskf4, const void *tbl16, int n_polys, cudaStream_t stream) {
    tb4::bsk_convert_kernel_v4<<<n_polys, 64, 0, stream>>>(bsk_std, reinterpret_cast<tb::cplx *>(bskf4),
                                                          reinterpret_cast<const tb::cplx *>(tbl16), n_polys);
    return cudaGetLastError();
}

}  // namespace tbk
