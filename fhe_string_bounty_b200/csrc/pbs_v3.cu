// pbs_v3.cu -- blind rotation, third generation: one 8-warp CTA per SM = 4 ciphertexts, Fourier bootstrapping
// key streamed ONCE per SM through an 80 KiB shared-memory ring by bulk asynchronous copies (TMA, cp.async.bulk +
// mbarrier), accumulator master copy in Tensor Memory (tcgen05.ld/st) so that shared memory has room for the ring.
//
// Same arithmetic as pbs.cu (bootstrap.rs:242-364, ggsw.rs:477-598, fft/mod.rs:197-326; see fft_core.cuh);
// what changes is where the data lives:
//   * ncu on the v2 kernel (profiles/r01_pbs_v2_summary.txt): 40 % of all warp stalls are long-scoreboard waits on the
//     Fourier-GGSW loads in the multiply-accumulate -- with 32 points per thread there are no registers to prefetch into,
//     and with 8 warps per SM nobody to hide an L2 round trip behind.  Asynchronous bulk copies need no registers.
//   * shared memory per SM: 8 x 16 KiB per-warp buffer (the polynomial for the rotated gather, then the transpose tile,
//     then the spectrum exchange) + 80 KiB ring + 16 KiB twiddle table = 224 KiB.  The accumulator's master copy
//     (each lane's own 64 coefficients) lives in TMEM: 32 lanes x 128 columns per warp, 256 columns per CTA.
//   * every ring piece (8 KiB = 4 FFT points x 2 x 2 GGSW polynomials) is consumed by all 8 warps; the last warp to
//     finish a piece re-arms its slot with the piece NSLOT ahead (1.25 iterations of lookahead), so the key crosses
//     L2 -> SM once per SM instead of once per ciphertext.
#include "kernels.h"
#include "fft_core.cuh"
#include "ring_helpers.cuh"

namespace tb3 {
using namespace tb;

// CTS = ciphertexts per CTA: 4 for throughput (8 warps share the ring); 1 for narrow tree levels (<= one ciphertext per
// SM: the two warps own a scheduler each, so an iteration's dependency chain runs without contention)
constexpr int PIECE_CPLX = 512;        // [out poly 2][sel 2][q 4][lane 32]
constexpr int PIECE_BYTES = PIECE_CPLX * 16;
constexpr int PIECES_PER_ITER = 8;
constexpr int NSLOT = 10;

template <int CTS>
struct Smem {
    static constexpr int WARPS = 2 * CTS;
    uint64_t mbuf[WARPS][kN];              // 16 KiB per warp
    cplx ring[NSLOT][PIECE_CPLX];          // 80 KiB
    cplx tbl[kM];                          // 16 KiB
    unsigned long long full_bar[NSLOT];
    unsigned int consumed[NSLOT];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem<4>) <= 227 * 1024, "shared memory budget");

using namespace tbr;
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// ---- TMEM (tcgen05) -------------------------------------------------------------------------------------------------
template <int TMEM_COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int TMEM_COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 columns (thread t <-> TMEM lane base+t, register i <-> column base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

__device__ __forceinline__ void warp_transpose(double (&v)[32], double *xb, int lane) {
#pragma unroll
    for (int r = 0; r < 32; ++r) xb[xpose_write_idx(lane, r)] = v[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = xb[xpose_read_idx(lane, r)];
    __syncwarp();
}

// Fourier key, v3 layout: [ggsw i][chunk 8][out poly c][sel: 0 = row c, 1 = row 1-c][q 4][lane 32]; point p = 4*chunk + q
__device__ __forceinline__ size_t bskf3_index(int i, int chunk, int c, int sel, int q) {
    return ((((size_t)(i * PIECES_PER_ITER + chunk) * 2 + c) * 2 + sel) * 4 + q) * 32;
}

template <int CTS>
__global__ void __launch_bounds__(64 * CTS, 1)
pbs_classic_kernel_v3(const uint64_t *__restrict__ lwe_small, const uint32_t *__restrict__ lut_idx, const uint64_t *__restrict__ luts,
                      const cplx *__restrict__ bskf3, const cplx *__restrict__ tbl_g, uint64_t *__restrict__ out,
                      const uint32_t *__restrict__ out_slot, int batch, int n, int base_log, int n_iters, int small_is_u16) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int WARPS = 2 * CTS, NTHREADS = 64 * CTS, TMEM_COLS = CTS > 2 ? 256 : 128;
    Smem<CTS> &sm = *reinterpret_cast<Smem<CTS> *>(smem_raw);
    const int W = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp -> (ciphertext, polynomial): the two warps of a ciphertext (coupled by pair barriers) sit on DIFFERENT
    // schedulers (warp id % 4), so each scheduler hosts warps of two different ciphertexts, which are started half an
    // iteration apart (see the stagger below): one is in an FP64 phase while the other is in a shared-memory phase.
    const int ctl = W >> 1, w = W & 1;
    const int ct_raw = blockIdx.x * CTS + ctl;
    const bool live = ct_raw < batch;
    const int ct = live ? ct_raw : batch - 1;        // ragged tail: recompute the last ciphertext, skip the store
    uint64_t *my = sm.mbuf[W];
    double *tile = reinterpret_cast<double *>(sm.mbuf[W]);
    cplx *myc = reinterpret_cast<cplx *>(sm.mbuf[W]);
    const cplx *othc = reinterpret_cast<const cplx *>(sm.mbuf[W ^ 1]);
    // input: n+1 u64 words, or (fused keyswitch + modulus switch) n+1 u16 values already switched to [0, 2N]
    const uint64_t *lwe = lwe_small + (size_t)ct * (n + 1);
    const uint16_t *lwe16 = reinterpret_cast<const uint16_t *>(lwe_small) + (size_t)ct * (n + 1);
    const int total_pieces = n_iters * PIECES_PER_ITER;

    // ---- one-time setup: twiddle table, barriers, TMEM, first ring fill ---------------------------------------------
    for (int i = threadIdx.x; i < kM; i += NTHREADS) sm.tbl[i] = tbl_g[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm.full_bar[s], 1); sm.consumed[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (W == 0) tmem_alloc<TMEM_COLS>(&sm.tmem_base);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_mine = sm.tmem_base + ((uint32_t)((W & 3) * 32) << 16) + (uint32_t)((W >> 2) * 128);   // lane quarter = warp id % 4
    if (threadIdx.x == 0) {
        const int first = total_pieces < NSLOT ? total_pieces : NSLOT;
        for (int g = 0; g < first; ++g) {
            mbar_expect_tx(&sm.full_bar[g], PIECE_BYTES);
            tma_load_1d(sm.ring[g], bskf3 + (size_t)g * PIECE_CPLX, PIECE_BYTES, &sm.full_bar[g]);
        }
    }

    // ---- acc <- LUT * X^(-b_hat): registers (own coefficients, as u64 bit patterns in re/im), TMEM, shared -------------------
    double re[32], im[32];
    {
        const uint32_t b_hat = (small_is_u16 ? (uint32_t)__ldg(lwe16 + n) : modulus_switch_2n(__ldg(lwe + n))) & (2 * kN - 1);
        const uint32_t a0 = (2 * kN - b_hat) & (2 * kN - 1);
        const uint64_t *lut = luts + ((size_t)(lut_idx ? lut_idx[ct] : 0) * 2 + w) * kN;
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            const int j = lane + 32 * m;
            int s0, s1; bool n0, n1;
            rot_src(j, a0, s0, n0);
            rot_src(j + kM, a0, s1, n1);
            uint64_t v0 = __ldg(lut + s0), v1 = __ldg(lut + s1);
            v0 = n0 ? (uint64_t)0 - v0 : v0;
            v1 = n1 ? (uint64_t)0 - v1 : v1;
            my[j] = v0; my[j + kM] = v1;
            re[m] = __longlong_as_double((long long)v0);
            im[m] = __longlong_as_double((long long)v1);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
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
    __syncwarp();

    // stagger: ciphertexts 2,3 start once ciphertexts 0,1 have finished their first forward FFT
    // stagger: ciphertext k starts when ciphertext 0 reaches the k-th quarter of its first iteration (after the gather,
    // after the forward FFT, after the MAC), so FP64 phases and shared-memory phases of different ciphertexts overlap
    if (CTS == 4 && ctl >= 1 && n_iters > 0) asm volatile("bar.sync %0, 128;" ::"r"(8 + ctl) : "memory");

    for (int i = 0; i < n_iters; ++i) {
        const uint32_t a = (small_is_u16 ? (uint32_t)__ldg(lwe16 + i) : modulus_switch_2n(__ldg(lwe + i))) & (2 * kN - 1);   // a == 0 is NOT skipped

        // ct1 = acc * X^a - acc, level-1 signed digit, folded (own coefficients come from the registers)
#pragma unroll
        for (int m0 = 0; m0 < 32; m0 += 8) {
#pragma unroll
            for (int m = m0; m < m0 + 8; ++m) {
                const int j = lane + 32 * m;
                const uint32_t s0 = ((uint32_t)j - a) & (2 * kN - 1);
                const uint32_t s1 = (s0 + kM) & (2 * kN - 1);
                uint64_t r0 = my[s0 & (kN - 1)], r1 = my[s1 & (kN - 1)];
                r0 = (s0 >= (uint32_t)kN) ? (uint64_t)0 - r0 : r0;
                r1 = (s1 >= (uint32_t)kN) ? (uint64_t)0 - r1 : r1;
                const uint64_t o0 = (uint64_t)__double_as_longlong(re[m]), o1 = (uint64_t)__double_as_longlong(im[m]);
                re[m] = (double)signed_digit_l1(r0 - o0, base_log);
                im[m] = (double)signed_digit_l1(r1 - o1, base_log);
            }
            TB_FENCE();
        }
        __syncwarp();   // everyone is done reading the polynomial: the buffer becomes the transpose tile
        if (CTS == 4 && i == 0 && ctl == 0) asm volatile("bar.arrive 9, 128;" ::: "memory");

        // forward FFT; the two radix-32 passes share ONE copy of the butterfly code (2-trip loop): the kernel body must stay
        // inside the instruction cache (ncu: no_instruction stall 0.08 -> 1.0 per instruction once it does not)
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0) pretwist_fwd(re, im);
            radix32_dif(re, im);
            if (pass == 0) {
                twiddle_fwd(re, im, [&](int idx) { return sm.tbl[idx]; }, lane);
                warp_transpose(re, tile, lane);
                warp_transpose(im, tile, lane);
            }
        }
        if (CTS == 4 && i == 0 && ctl == 0) asm volatile("bar.arrive 10, 128;" ::: "memory");

        // spectrum exchange between the two warps of the ciphertext (whole polynomial at once: 16 KiB buffer)
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            cplx f; f.x = re[p]; f.y = im[p];
            myc[p * 32 + lane] = f;
        }
        pair_barrier(1 + ctl);

        // out_fft[w] = F_w * G[w][w] + F_{1-w} * G[1-w][w], GGSW pieces from the ring.
        // (1) probe the 8 piece barriers back to back (they completed long ago: the ring runs 1.25 iterations ahead), only
        //     falling back to a blocking wait for a piece that is really missing;
        // (2) per chunk: 12 independent 16-byte shared loads, then 32 FP64 instructions -- one exposed latency per chunk;
        // (3) release each slot right after its chunk (its values have been consumed by then, so no fence is needed for load
        //     completion).
        {
            const int g0 = i * PIECES_PER_ITER;
            uint32_t ready = 0;
#pragma unroll
            for (int c = 0; c < PIECES_PER_ITER; ++c) {
                const int g = g0 + c;
                ready |= (mbar_try_wait(&sm.full_bar[g % NSLOT], (uint32_t)(g / NSLOT) & 1u) ? 1u : 0u) << c;
            }
#pragma unroll
            for (int c = 0; c < PIECES_PER_ITER; ++c) {
                const int g = g0 + c;
                const int slot = g % NSLOT;
                if (!((ready >> c) & 1u)) mbar_wait(&sm.full_bar[slot], (uint32_t)(g / NSLOT) & 1u);
                const cplx *pc = sm.ring[slot] + (w * 2) * 4 * 32 + lane;
                cplx ga[4], gb[4], fo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ga[q] = pc[q * 32];
                    gb[q] = pc[(4 + q) * 32];
                    fo[q] = othc[(c * 4 + q) * 32 + lane];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int p = c * 4 + q;
                    const double fr = re[p], fi = im[p];
                    double orr = DMUL(fr, ga[q].x);
                    orr = DFMA(-fi, ga[q].y, orr);
                    orr = DFMA(fo[q].x, gb[q].x, orr);
                    orr = DFMA(-fo[q].y, gb[q].y, orr);
                    double oi = DMUL(fr, ga[q].y);
                    oi = DFMA(fi, ga[q].x, oi);
                    oi = DFMA(fo[q].x, gb[q].y, oi);
                    oi = DFMA(fo[q].y, gb[q].x, oi);
                    re[p] = orr; im[p] = oi;
                }
                // release the slot as soon as this warp is done with it (progressive release keeps the ring full for the
                // ciphertext group that runs half an iteration ahead); the last of the 8 warps re-arms it NSLOT pieces ahead
                __syncwarp();
                if (lane == 0) {
                    const unsigned int old = atomicAdd(&sm.consumed[slot], 1u);
                    if (old == WARPS - 1) {
                        sm.consumed[slot] = 0;
                        const int g2 = g + NSLOT;
                        if (g2 < total_pieces) {
                            __threadfence_block();
                            fence_proxy_async();
                            mbar_expect_tx(&sm.full_bar[slot], PIECE_BYTES);
                            tma_load_1d(sm.ring[slot], bskf3 + (size_t)g2 * PIECE_CPLX, PIECE_BYTES, &sm.full_bar[slot]);
                        }
                    }
                }
            }
        }
        pair_barrier(1 + ctl);   // the partner has read my spectrum: the buffer is the transpose tile again
        if (CTS == 4 && i == 0 && ctl == 0) asm volatile("bar.arrive 11, 128;" ::: "memory");

#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            radix32_dit_inv(re, im);
            if (pass == 0) {
                warp_transpose(re, tile, lane);
                warp_transpose(im, tile, lane);
                twiddle_inv(re, im, [&](int idx) { return sm.tbl[idx]; }, lane);
            } else {
                posttwist_inv(re, im);
            }
        }

        // acc += from_torus(.): master copy in TMEM, new values to registers (next gather's "own") and shared (next rotation)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t v[16];
            tmem_ld16(tmem_mine + 16 * k, v);
            tmem_wait_ld();
#pragma unroll
            for (int mm = 0; mm < 4; ++mm) {
                const int m = 4 * k + mm, j = lane + 32 * m;
                uint64_t o0 = ((uint64_t)v[4 * mm + 1] << 32) | v[4 * mm];
                uint64_t o1 = ((uint64_t)v[4 * mm + 3] << 32) | v[4 * mm + 2];
                o0 += from_torus_f64(re[m]);
                o1 += from_torus_f64(im[m]);
                v[4 * mm] = (uint32_t)o0; v[4 * mm + 1] = (uint32_t)(o0 >> 32);
                v[4 * mm + 2] = (uint32_t)o1; v[4 * mm + 3] = (uint32_t)(o1 >> 32);
                my[j] = o0; my[j + kM] = o1;
                re[m] = __longlong_as_double((long long)o0);
                im[m] = __longlong_as_double((long long)o1);
            }
            tmem_st16(tmem_mine + 16 * k, v);
        }
        tmem_wait_st();
        __syncwarp();
    }

    // sample extraction (coefficient 0) straight from the registers: out[0] = A[0], out[N-j] = -A[j]; body = B[0]
    if (live) {
        uint64_t *o = out + (size_t)(out_slot ? out_slot[ct] : ct) * (kN + 1);
        if (w == 0) {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const int j = lane + 32 * m;
                const uint64_t v0 = (uint64_t)__double_as_longlong(re[m]), v1 = (uint64_t)__double_as_longlong(im[m]);
                if (j == 0) o[0] = v0; else o[kN - j] = (uint64_t)0 - v0;
                o[kN - (j + kM)] = (uint64_t)0 - v1;
            }
        } else if (lane == 0) {
            o[kN] = (uint64_t)__double_as_longlong(re[0]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (W == 0) tmem_dealloc<TMEM_COLS>(sm.tmem_base);
}

// std -> Fourier key in the v3 ring layout (one warp per polynomial; same forward transform as the kernel above)
__global__ void __launch_bounds__(32)
bsk_convert_kernel_v3(const uint64_t *__restrict__ bsk_std, cplx *__restrict__ bskf3, const cplx *__restrict__ tbl, int n_polys) {
    __shared__ double xb[kXposeWords];
    const int qd = blockIdx.x, lane = threadIdx.x;
    if (qd >= n_polys) return;
    const int i = qd >> 2, r = (qd >> 1) & 1, c = qd & 1;   // std layout [i][level 1][row r][col c][N]
    const uint64_t *src = bsk_std + (size_t)qd * kN;
    const double scale = 5.293955920339377e-23;              // 2^-74
    double re[32], im[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) {
        const int j = lane + 32 * m;
        re[m] = DMUL((double)(long long)src[j], scale);
        im[m] = DMUL((double)(long long)src[j + kM], scale);
    }
    pretwist_fwd(re, im);
    radix32_dif(re, im);
    twiddle_fwd(re, im, [&](int idx) { return __ldg(tbl + idx); }, lane);
    warp_transpose(re, xb, lane);
    warp_transpose(im, xb, lane);
    radix32_dif(re, im);
    const int sel = (r == c) ? 0 : 1;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        cplx v; v.x = re[p]; v.y = im[p];
        bskf3[bskf3_index(i, p >> 2, c, sel, p & 3) + lane] = v;
    }
}

}  // namespace tb3

namespace tbk {

cudaError_t pbs_v3_configure() {
    cudaError_t e = cudaFuncSetAttribute(tb3::pbs_classic_kernel_v3<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tb3::Smem<4>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tb3::pbs_classic_kernel_v3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tb3::Smem<2>));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tb3::pbs_classic_kernel_v3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tb3::Smem<1>));
}

cudaError_t launch_pbs_classic_v3(const uint64_t *lwe_small, const uint32_t *lut_idx, const uint64_t *luts, const void *bskf3,
                                  const void *tbl, uint64_t *out, const uint32_t *out_slot, int batch, int n, int base_log,
                                  int n_iters, int small_is_u16, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (batch <= sms) {   // narrow level: one ciphertext per SM, latency-oriented instance
        tb3::pbs_classic_kernel_v3<1><<<batch, 64, sizeof(tb3::Smem<1>), stream>>>(
            lwe_small, lut_idx, luts, reinterpret_cast<const tb::cplx *>(bskf3), reinterpret_cast<const tb::cplx *>(tbl), out, out_slot,
            batch, n, base_log, n_iters, small_is_u16);
        return cudaGetLastError();
    }
    if (batch <= 2 * sms) {   // up to two ciphertexts per SM: still one warp per scheduler
        tb3::pbs_classic_kernel_v3<2><<<(batch + 1) / 2, 128, sizeof(tb3::Smem<2>), stream>>>(
            lwe_small, lut_idx, luts, reinterpret_cast<const tb::cplx *>(bskf3), reinterpret_cast<const tb::cplx *>(tbl), out, out_slot,
            batch, n, base_log, n_iters, small_is_u16);
        return cudaGetLastError();
    }
    const int grid = (batch + 3) / 4;
    tb3::pbs_classic_kernel_v3<4><<<grid, 256, sizeof(tb3::Smem<4>), stream>>>(
        lwe_small, lut_idx, luts, reinterpret_cast<const tb::cplx *>(bskf3), reinterpret_cast<const tb::cplx *>(tbl), out, out_slot,
        batch, n, base_log, n_iters, small_is_u16);
    return cudaGetLastError();
}

cudaError_t launch_bsk_convert_v3(const uint64_t *bsk_std, void *bskf3, const void *tbl, int n_polys, cudaStream_t stream) {
    tb3::bsk_convert_kernel_v3<<<n_polys, 32, 0, stream>>>(bsk_std, reinterpret_cast<tb::cplx *>(bskf3),
                                                          reinterpret_cast<const tb::cplx *>(tbl), n_polys);
    return cudaGetLastError();
}

}  // namespace tbk
