// pbs8_common.cuh -- device helpers shared by the 8-points-per-thread blind-rotation kernels (pbs_v8.cu, pbs_multibit_v8.cu): the
// 128-thread FFT of fft8_core.cuh over a shared-memory tile (exchanges A and B) and register shuffles (exchange C), and the three
// places its 24 per-thread twiddles can come from (registers, Tensor Memory, the global table).
#pragma once
#include "fft8_core.cuh"
#include "pbs16_common.cuh"

namespace tb8c {
using namespace tb;          // cplx, kN, kM, integer helpers of the blind rotation
using namespace tbr;         // mbarrier / bulk-copy helpers
using namespace tb8;         // the 128-thread FFT
using tb16k::bar_sync;
using tb16k::BlockSync;
using tb16k::cplx_from_words;
using tb16k::tmem_alloc;
using tb16k::tmem_dealloc;
using tb16k::tmem_ld16;
using tb16k::tmem_st16;
using tb16k::tmem_wait_ld;
using tb16k::tmem_wait_st;

struct PolySync128 {
    int id;
    __device__ __forceinline__ void operator()() const { bar_sync(id, 128); }
};

// the 24 per-thread twiddles (fft8_core.cuh: T1, T2, T3) from this thread's TMEM lane (96 columns) or from the global table
struct TmemTw8 {
    uint32_t col;
    __device__ __forceinline__ void load(int block, cplx (&tw)[8]) const {
        uint32_t v[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            tmem_ld16(col + 32 * block + 16 * h, v);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) tw[4 * h + q] = cplx_from_words(v, q);
        }
    }
};
// one ciphertext per CTA leaves 255 registers per thread: all 24 twiddles simply stay in registers
struct RegTw8 {
    const cplx (&t)[24];
    __device__ __forceinline__ void load(int block, cplx (&tw)[8]) const {
#pragma unroll
        for (int p = 0; p < 8; ++p) tw[p] = t[8 * block + p];
    }
};
struct GlobalTw8 {
    const cplx *row;     // table + 24 * T
    __device__ __forceinline__ void load(int block, cplx (&tw)[8]) const {
#pragma unroll
        for (int p = 0; p < 8; ++p) tw[p] = __ldg(row + 8 * block + p);
    }
};

template <class F>
__device__ __forceinline__ void st8(cplx *base, const double (&re)[8], const double (&im)[8], F off) {
#pragma unroll
    for (int p = 0; p < 8; ++p) { cplx v; v.x = re[p]; v.y = im[p]; base[off(p)] = v; }
}
template <class F>
__device__ __forceinline__ void ld8(const cplx *base, double (&re)[8], double (&im)[8], F off) {
#pragma unroll
    for (int p = 0; p < 8; ++p) { const cplx v = base[off(p)]; re[p] = v.x; im[p] = v.y; }
}

// Exchange C between the two lanes of a pair through ONE shuffle per word instead of the tile (xc_* in fft8_core.cuh, which the CPU
// mirror checks and which costs as many shared-memory wavefronts as exchanges A and B together because it is 2-way bank conflicted).
// Lane bit b = c' (before) = e' (after).  Forward: register pe + 4*cl -> c + 4*pe0; a lane keeps its pe1 = b values and receives the
// partner's.  The kept value goes to position c = cl, the received one to c = cl + 2 on BOTH lanes; on the b = 1 lane that is the wrong
// way round (c and c + 2 swapped), which a radix-4 turns into a factor (-1)^f on its outputs -- fixed by negating the odd-f outputs.
__device__ __forceinline__ void flip_odd_f(double (&re)[8], double (&im)[8], int b) {
    if (b) {
#pragma unroll
        for (int g = 0; g < 8; g += 4) { re[g + 2] = -re[g + 2]; im[g + 2] = -im[g + 2]; re[g + 3] = -re[g + 3]; im[g + 3] = -im[g + 3]; }
    }
}
__device__ __forceinline__ void exchange_c_fwd(double (&re)[8], double (&im)[8], int b) {
    double ore[8], oim[8];
#pragma unroll
    for (int cl = 0; cl < 2; ++cl)
#pragma unroll
        for (int pe0 = 0; pe0 < 2; ++pe0) {
            const int i0 = pe0 + 4 * cl, i1 = i0 + 2;                   // pe1 = 0, 1
            const double kr = b ? re[i1] : re[i0], ki = b ? im[i1] : im[i0];
            const double sr = b ? re[i0] : re[i1], si = b ? im[i0] : im[i1];
            ore[cl + 4 * pe0] = kr; oim[cl + 4 * pe0] = ki;
            ore[cl + 2 + 4 * pe0] = __shfl_xor_sync(0xffffffffu, sr, 1);
            oim[cl + 2 + 4 * pe0] = __shfl_xor_sync(0xffffffffu, si, 1);
        }
#pragma unroll
    for (int r = 0; r < 8; ++r) { re[r] = ore[r]; im[r] = oim[r]; }
}
// inverse: register c + 4*pe0 (already swapped on the b = 1 lane by negating the odd-f INPUTS of the inverse radix-4) -> pe + 4*cl
__device__ __forceinline__ void exchange_c_inv(double (&re)[8], double (&im)[8], int b) {
    double ore[8], oim[8];
#pragma unroll
    for (int cl = 0; cl < 2; ++cl)
#pragma unroll
        for (int pe0 = 0; pe0 < 2; ++pe0) {
            const double kr = re[cl + 4 * pe0], ki = im[cl + 4 * pe0];
            const double rr = __shfl_xor_sync(0xffffffffu, re[cl + 2 + 4 * pe0], 1), ri = __shfl_xor_sync(0xffffffffu, im[cl + 2 + 4 * pe0], 1);
            const int i0 = pe0 + 4 * cl, i1 = i0 + 2;
            ore[i0] = b ? rr : kr; oim[i0] = b ? ri : ki;
            ore[i1] = b ? kr : rr; oim[i1] = b ? ki : ri;
        }
#pragma unroll
    for (int r = 0; r < 8; ++r) { re[r] = ore[r]; im[r] = oim[r]; }
}

// ---- two-SM cluster instances (pbs_classic_kernel_v8x2, pbs_multibit_kernel_v8x2): mbarrier hand-offs inside and across the CTAs ----
__device__ __forceinline__ void mbar_arrive(void *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(void *bar, uint32_t parity) {     // acquire at cluster scope: the partner's remote stores are visible after it
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) __trap();
    }
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double x, double y) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
// remote store that signals the destination CTA's mbarrier with the bytes it delivered (complete_tx): no fence, no separate arrival
__device__ __forceinline__ void st_async_cluster(uint32_t addr, double x, double y, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(addr), "d"(x), "d"(y), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// forward: on entry the tile may still be read by other threads (the first sync covers that); on exit thread t holds register r =
// frequency freq_of8(t, r) and nobody but t touches t's exchange-C reader slots.
template <class Tw, class Sync>
__device__ __forceinline__ void fft8_fwd(double (&re)[8], double (&im)[8], cplx *tile, const Tw &twd, int T, Sync sync) {
    cplx tw[8];
    pretwist8_fwd(re, im);
    radix8_dif(re, im);
    twd.load(0, tw);
    twiddle8<false>(re, im, tw, 0);
    sync();
    st8(tile + xa_wbase(T), re, im, [](int p) { return xa_woff(p); });
    sync();
    ld8(tile + xa_rbase(T), re, im, [](int p) { return xa_roff(p); });
    radix8_dif(re, im);
    twd.load(1, tw);
    twiddle8<false>(re, im, tw, 1);
    __syncwarp();     // from here on everything stays inside this half-warp's region of the tile
    st8(tile + xb_wbase(T), re, im, [](int p) { return xb_woff(p); });
    __syncwarp();
    ld8(tile + xb_rbase(T), re, im, [](int p) { return xb_roff(p); });
    radix4x2_dif(re, im);
    twd.load(2, tw);
    twiddle8<false>(re, im, tw, 0);
    exchange_c_fwd(re, im, T & 1);
    radix4x2_dif(re, im);
    flip_odd_f(re, im, T & 1);
    __syncwarp();     // the other lanes of the half-warp are done reading exchange B: the C-reader slots may be reused for the spectrum
}

// inverse (scaled by 1024): on entry nobody else may be reading this thread's exchange-C reader slots; on exit the tile may still be
// read by other threads.
template <class Tw, class Sync>
__device__ __forceinline__ void fft8_inv(double (&re)[8], double (&im)[8], cplx *tile, const Tw &twd, int T, Sync sync) {
    cplx tw[8];
    flip_odd_f(re, im, T & 1);
    radix4x2_dit_inv(re, im);
    exchange_c_inv(re, im, T & 1);
    twd.load(2, tw);
    twiddle8<true>(re, im, tw, 0);
    radix4x2_dit_inv(re, im);
    __syncwarp();
    st8(tile + xb_rbase(T), re, im, [](int p) { return xb_roff(p); });
    __syncwarp();
    ld8(tile + xb_wbase(T), re, im, [](int p) { return xb_woff(p); });
    twd.load(1, tw);
    twiddle8<true>(re, im, tw, 1);
    radix8_dit_inv(re, im);
    __syncwarp();
    st8(tile + xa_rbase(T), re, im, [](int p) { return xa_roff(p); });
    sync();
    ld8(tile + xa_wbase(T), re, im, [](int p) { return xa_woff(p); });
    twd.load(0, tw);
    twiddle8<true>(re, im, tw, 0);
    radix8_dit_inv(re, im);
    posttwist8_inv(re, im);
}

}  // namespace tb8c
