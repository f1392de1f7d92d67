// pbs_multibit.cu -- multi-bit programmable bootstrap (grouping factor g = 3) for sm_100a.
//
// Replaces core_crypto/algorithms/lwe_multi_bit_programmable_bootstrapping.rs:
//   prepare_multi_bit_ggsw_mem_optimized  :18-84    (G = G_0 + sum_{j=1..7} G_j * X^{ms(sum of selected a_t)})
//   multi_bit_blind_rotate_assign         :295-546  / deterministic variant :548-  (acc_new = G (x) acc_old, dst zeroed :503)
//   incomplete_monomial_forward_as_integer fft64/math/fft/mod.rs:408-445, update_with_fmadd_factor fft64/crypto/ggsw.rs:699-754
// and the std->Fourier conversion of entities/lwe_multi_bit_bootstrap_key.rs:11-62,318.
// The CPU runs 7 producer threads building G and 1 consumer (arrival order => non-deterministic rounding unless
// deterministic_execution); here every ciphertext combines its own G on the fly in a fixed order (j = 1..7), so the
// result is deterministic run to run.
//
// B200 mapping: because the monomial rotation X^{deg} is applied in the Fourier domain (point-wise factor
// zeta_k^{deg}), a blind-rotation step needs NO coefficient-domain rotation: the accumulator never leaves the lanes'
// registers (each lane keeps its 64 coefficients as u64 bit patterns in the FFT register file), there is no shared
// accumulator, no gather, no TMEM.  Per step: digits of own coefficients -> forward FFT -> spectrum exchange ->
// for each point  Gc = G_0 + sum_j G_j * M_j  (M_j[k] = A_{lane,j} * W32^{deg_j * brev5(p)}: one root-table load per
// lane and j, the register-dependent part from constant memory) -> 2x2 MAC -> inverse FFT -> round to torus (replace).
// The grouped Fourier key (512 KiB per step, 155 MB in total, larger than L2) streams once per SM through a 128 KiB
// shared-memory ring of bulk asynchronous copies.
#include "kernels.h"
#include "fft_core.cuh"
#include "ring_helpers.cuh"

namespace tbm {
using namespace tb;
using namespace tbr;

constexpr int CTS = 4;
constexpr int WARPS = 2 * CTS;
constexpr int NTHREADS = 32 * WARPS;
constexpr int GF = 3;                          // grouping factor
constexpr int NGGSW = 1 << GF;                 // GGSWs per group
constexpr int PIECE_CPLX = 1024;               // [jl 2][out poly 2][sel 2][q 4][lane 32]
constexpr int PIECE_BYTES = PIECE_CPLX * 16;   // 16 KiB
constexpr int PIECES_PER_CHUNK = NGGSW / 2;    // 4
constexpr int CHUNKS = 8;                      // 4 FFT points each
constexpr int PIECES_PER_ITER = CHUNKS * PIECES_PER_CHUNK;   // 32 = 512 KiB per group
constexpr int NSLOT = 8;                       // 128 KiB ring

__constant__ double c_w32[32][2];              // exp(-2*pi*i*e/32)

struct Smem {
    double xb[WARPS][kXposeWords];             // transpose tile / half-spectrum exchange (8448 B each)
    cplx ring[NSLOT][PIECE_CPLX];
    cplx tbl[kM];
    unsigned long long full_bar[NSLOT];
    unsigned int consumed[NSLOT];
};
static_assert(sizeof(Smem) <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ void warp_transpose(double (&v)[32], double *xb, int lane) {
#pragma unroll
    for (int r = 0; r < 32; ++r) xb[xpose_write_idx(lane, r)] = v[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = xb[xpose_read_idx(lane, r)];
    __syncwarp();
}

// Fourier key layout: piece g = (group*8 + chunk)*4 + (j >> 1); inside [j & 1][out poly c][sel][q][lane]
__device__ __forceinline__ size_t bskm_index(int grp, int j, int chunk, int c, int sel, int q) {
    return ((((((size_t)(grp * CHUNKS + chunk) * PIECES_PER_CHUNK + (j >> 1)) * 2 + (j & 1)) * 2 + c) * 2 + sel) * 4 + q) * 32;
}

__global__ void __launch_bounds__(NTHREADS, 1)
pbs_multibit_kernel(const uint64_t *__restrict__ lwe_small, const uint32_t *__restrict__ lut_idx, const uint64_t *__restrict__ luts,
                    const cplx *__restrict__ bskm, const cplx *__restrict__ tbl_g, const cplx *__restrict__ roots,   // roots[e] = exp(i*pi*e/2048), e < 4096
                    uint64_t *__restrict__ out, const uint32_t *__restrict__ out_slot, int batch, int n, int base_log,
                    int n_groups) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int W = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctl = W >> 1, w = W & 1;
    const int ct_raw = blockIdx.x * CTS + ctl;
    const bool live = ct_raw < batch;
    const int ct = live ? ct_raw : batch - 1;
    double *tile = sm.xb[W];
    cplx *myc = reinterpret_cast<cplx *>(sm.xb[W]);
    const cplx *othc = reinterpret_cast<const cplx *>(sm.xb[W ^ 1]);
    const uint64_t *lwe = lwe_small + (size_t)ct * (n + 1);
    const int total_pieces = n_groups * PIECES_PER_ITER;

    for (int i = threadIdx.x; i < kM; i += NTHREADS) sm.tbl[i] = tbl_g[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm.full_bar[s], 1); sm.consumed[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int first = total_pieces < NSLOT ? total_pieces : NSLOT;
        for (int g = 0; g < first; ++g) {
            mbar_expect_tx(&sm.full_bar[g], PIECE_BYTES);
            tma_load_1d(sm.ring[g], bskm + (size_t)g * PIECE_CPLX, PIECE_BYTES, &sm.full_bar[g]);
        }
    }

    // acc <- LUT * X^(-b_hat) (lwe_multi_bit_programmable_bootstrapping.rs:373-391), own coefficients only
    double re[32], im[32];
    {
        const uint32_t b_hat = modulus_switch_2n(__ldg(lwe + n)) & (2 * kN - 1);
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
            re[m] = __longlong_as_double((long long)v0);
            im[m] = __longlong_as_double((long long)v1);
        }
    }
    const int rot_t = (1 - 4 * brev5(lane)) & (2 * kN - 1);   // exponent of the lane-dependent part of zeta_k = w^(1 - 4k)

    for (int grp = 0; grp < n_groups; ++grp) {
        // monomial degrees of the 7 non-constant GGSWs (:44-62): bit (g-1-t) of j selects mask element t; modulus switch of the SUM
        const uint64_t a0v = __ldg(lwe + GF * grp), a1v = __ldg(lwe + GF * grp + 1), a2v = __ldg(lwe + GF * grp + 2);
        uint32_t deg[NGGSW];
        cplx A[NGGSW];
#pragma unroll
        for (int j = 1; j < NGGSW; ++j) {
            const uint64_t s = ((j & 4) ? a0v : 0) + ((j & 2) ? a1v : 0) + ((j & 1) ? a2v : 0);
            deg[j] = modulus_switch_2n(s) & (2 * kN - 1);
            A[j] = __ldg(roots + ((deg[j] * (uint32_t)rot_t) & (2 * kN - 1)));   // w^(deg * (1 - 4*brev5(lane)))
        }

        // decomposition of the accumulator itself (ggsw.rs:515-533 on src = acc_old), folded
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            re[m] = (double)signed_digit_l1((uint64_t)__double_as_longlong(re[m]), base_log);
            im[m] = (double)signed_digit_l1((uint64_t)__double_as_longlong(im[m]), base_log);
        }

        // the two radix-32 passes share one copy of the butterfly code: the kernel must fit the instruction cache
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

#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int pp = 0; pp < 16; ++pp) {
                cplx f; f.x = re[half * 16 + pp]; f.y = im[half * 16 + pp];
                myc[pp * 32 + lane] = f;
            }
            pair_barrier(1 + ctl);
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
                const int chunk = half * 4 + c4;
                cplx Ga[4], Gb[4];
#pragma unroll
                for (int pc = 0; pc < PIECES_PER_CHUNK; ++pc) {
                    const int g = (grp * CHUNKS + chunk) * PIECES_PER_CHUNK + pc;
                    const int slot = g % NSLOT;
                    mbar_wait(&sm.full_bar[slot], (uint32_t)(g / NSLOT) & 1u);
#pragma unroll
                    for (int jl = 0; jl < 2; ++jl) {
                        const int j = pc * 2 + jl;
                        const cplx *base = sm.ring[slot] + ((jl * 2 + w) * 2) * 4 * 32 + lane;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const cplx ga = base[q * 32], gb = base[(4 + q) * 32];
                            if (j == 0) {
                                Ga[q] = ga; Gb[q] = gb;
                            } else {
                                // M = A_j * W32^(deg_j * brev5(p)): monomial spectrum at this lane's point p (fft/mod.rs:413-444)
                                const int p = chunk * 4 + q;
                                const uint32_t e = (deg[j] * (uint32_t)brev5(p)) & 31u;
                                const double br = c_w32[e][0], bi = c_w32[e][1];
                                const double mr = DFMA(A[j].x, br, -DMUL(A[j].y, bi));
                                const double mi = DFMA(A[j].x, bi, DMUL(A[j].y, br));
                                Ga[q].x = DFMA(ga.x, mr, DFMA(-ga.y, mi, Ga[q].x));
                                Ga[q].y = DFMA(ga.x, mi, DFMA(ga.y, mr, Ga[q].y));
                                Gb[q].x = DFMA(gb.x, mr, DFMA(-gb.y, mi, Gb[q].x));
                                Gb[q].y = DFMA(gb.x, mi, DFMA(gb.y, mr, Gb[q].y));
                            }
                        }
                    }
                    // all lanes have consumed the piece (its values fed the arithmetic above): release the slot
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
                                tma_load_1d(sm.ring[slot], bskm + (size_t)g2 * PIECE_CPLX, PIECE_BYTES, &sm.full_bar[slot]);
                            }
                        }
                    }
                }
                // out_fft[w] = F_w * Gc[w][w] + F_{1-w} * Gc[1-w][w]   (register arrays need compile-time indices: switch on c4)
                cplx fo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) fo[q] = othc[(c4 * 4 + q) * 32 + lane];
                switch (c4) {
#define TBM_MAC(C)                                                                                   \
    case C: {                                                                                        \
        _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                               \
            const int p = (half * 4 + (C)) * 4 + q;                                                  \
            const double fr = re[p], fi = im[p];                                                     \
            double orr = DMUL(fr, Ga[q].x);                                                          \
            orr = DFMA(-fi, Ga[q].y, orr);                                                           \
            orr = DFMA(fo[q].x, Gb[q].x, orr);                                                       \
            orr = DFMA(-fo[q].y, Gb[q].y, orr);                                                      \
            double oi = DMUL(fr, Ga[q].y);                                                           \
            oi = DFMA(fi, Ga[q].x, oi);                                                              \
            oi = DFMA(fo[q].x, Gb[q].y, oi);                                                         \
            oi = DFMA(fo[q].y, Gb[q].x, oi);                                                         \
            re[p] = orr; im[p] = oi;                                                                 \
        }                                                                                            \
    } break;
                    TBM_MAC(0) TBM_MAC(1) TBM_MAC(2) TBM_MAC(3)
#undef TBM_MAC
                }
            }
            pair_barrier(1 + ctl);
        }

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

        // dst = 0; dst += G (x) src  (:503): the accumulator is REPLACED by the rounded product
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            re[m] = __longlong_as_double((long long)from_torus_f64(re[m]));
            im[m] = __longlong_as_double((long long)from_torus_f64(im[m]));
        }
    }

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
}

// std multi-bit key [group][j 8][level 1][row r][col c][N] (entities/lwe_multi_bit_bootstrap_key.rs:11-62) -> ring layout
__global__ void __launch_bounds__(32)
bsk_convert_multibit_kernel(const uint64_t *__restrict__ bsk_std, cplx *__restrict__ bskm, const cplx *__restrict__ tbl, int n_polys) {
    __shared__ double xb[kXposeWords];
    const int qd = blockIdx.x, lane = threadIdx.x;
    if (qd >= n_polys) return;
    const int c = qd & 1, r = (qd >> 1) & 1, j = (qd >> 2) & 7, grp = qd >> 5;
    const uint64_t *src = bsk_std + (size_t)qd * kN;
    const double scale = 5.293955920339377e-23;   // 2^-74
    double re[32], im[32];
#pragma unroll
    for (int m = 0; m < 32; ++m) {
        const int jj = lane + 32 * m;
        re[m] = DMUL((double)(long long)src[jj], scale);
        im[m] = DMUL((double)(long long)src[jj + kM], scale);
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
        bskm[bskm_index(grp, j, p >> 2, c, sel, p & 3) + lane] = v;
    }
}

}  // namespace tbm

namespace tbk {

cudaError_t pbs_multibit_configure() {
    double h[32][2];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int e = 0; e < 32; ++e) {
        h[e][0] = (double)cosl(-2.0L * pi * e / 32.0L);
        h[e][1] = (double)sinl(-2.0L * pi * e / 32.0L);
    }
    cudaError_t err = cudaMemcpyToSymbol(tbm::c_w32, h, sizeof(h));
    if (err != cudaSuccess) return err;
    return cudaFuncSetAttribute(tbm::pbs_multibit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tbm::Smem));
}

cudaError_t launch_pbs_multibit(const uint64_t *lwe_small, const uint32_t *lut_idx, const uint64_t *luts, const void *bskm,
                                const void *tbl, const void *roots, uint64_t *out, const uint32_t *out_slot, int batch, int n,
                                int base_log, int n_groups, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    const int grid = (batch + tbm::CTS - 1) / tbm::CTS;
    tbm::pbs_multibit_kernel<<<grid, tbm::NTHREADS, sizeof(tbm::Smem), stream>>>(
        lwe_small, lut_idx, luts, reinterpret_cast<const tb::cplx *>(bskm), reinterpret_cast<const tb::cplx *>(tbl),
        reinterpret_cast<const tb::cplx *>(roots), out, out_slot, batch, n, base_log, n_groups);
    return cudaGetLastError();
}

cudaError_t launch_bsk_convert_multibit(const uint64_t *bsk_std, void *bskm, const void *tbl, int n_polys, cudaStream_t stream) {
    tbm::bsk_convert_multibit_kernel<<<n_polys, 32, 0, stream>>>(bsk_std, reinterpret_cast<tb::cplx *>(bskm),
                                                                reinterpret_cast<const tb::cplx *>(tbl), n_polys);
    return cudaGetLastError();
}

}  // namespace tbk
