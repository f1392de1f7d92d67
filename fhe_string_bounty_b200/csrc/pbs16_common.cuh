// pbs16_common.cuh -- device helpers shared by the 16-points-per-thread blind-rotation kernels (pbs_v4.cu, pbs_multibit_v4.cu):
// named barriers, Tensor Memory access, twiddles kept in TMEM, and the 64-thread FFT of fft16_core.cuh over a shared-memory tile.
#pragma once
#include "fft16_core.cuh"
#include "ring_helpers.cuh"

namespace tb16k {
using namespace tb;
using namespace tb16;
using namespace tbr;

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

struct PolySync {      // named barrier over the 64 threads of one polynomial
    int id;
    __device__ __forceinline__ void operator()() const { bar_sync(id, 64); }
};
struct CtSync {        // named barrier over the 128 threads of one ciphertext (both polynomials)
    int id;
    __device__ __forceinline__ void operator()() const { bar_sync(id, 128); }
};
struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

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

// ---- twiddles from Tensor Memory: each thread keeps its own 16 inter-pass twiddles T1[p][T] (64 columns) and its three pass-2
// twiddles (12 columns) in its TMEM lane, so the blind-rotation loop reads them with tcgen05.ld instead of through the LSU pipe.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ cplx cplx_from_words(const uint32_t (&v)[16], int q) {
    cplx w;
    w.x = __hiloint2double((int)v[4 * q + 1], (int)v[4 * q]);
    w.y = __hiloint2double((int)v[4 * q + 3], (int)v[4 * q + 2]);
    return w;
}
struct TmemTwiddles {
    uint32_t col;     // TMEM address of this thread's first twiddle column
    template <bool INV>
    __device__ __forceinline__ void apply16(double (&re)[16], double (&im)[16]) const {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v[16];
            tmem_ld16(col + 16 * k, v);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const cplx w = cplx_from_words(v, q);
                const int p = 4 * k + q;
                const double a = re[p], b = im[p];
                if (!INV) {
                    re[p] = DFMA(a, w.x, -DMUL(b, w.y));
                    im[p] = DFMA(b, w.x, DMUL(a, w.y));
                } else {
                    re[p] = DFMA(a, w.x, DMUL(b, w.y));
                    im[p] = DFMA(b, w.x, -DMUL(a, w.y));
                }
            }
        }
    }
    __device__ __forceinline__ void load3(cplx (&tw)[3]) const {
        uint32_t v[16];
        tmem_ld16(col + 64, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 3; ++q) tw[q] = cplx_from_words(v, q);
    }
};
// the same twiddles straight from the global tables (key conversion kernel): identical values, identical arithmetic
struct GlobalTwiddles {
    const cplx *tbl16;
    int T;
    template <bool INV>
    __device__ __forceinline__ void apply16(double (&re)[16], double (&im)[16]) const {
        if (!INV) twiddle16_fwd(re, im, [&](int p) { return __ldg(tbl16 + p * 64 + T); });
        else twiddle16_inv(re, im, [&](int p) { return __ldg(tbl16 + p * 64 + T); });
    }
    __device__ __forceinline__ void load3(cplx (&tw)[3]) const {
#pragma unroll
        for (int q = 0; q < 3; ++q) tw[q] = __ldg(tbl16 + kM + 16 * q + (T & 15));
    }
};

// ---- the 64-thread FFT: registers <-> tile exchanges (fft16_core.cuh), `sync` = barrier over the polynomial's 64 threads -----------
// forward: on entry the tile may still be read by the other threads (the first sync covers that); on exit thread T2 holds
// register pv = frequency freq_of16(T2, pv) and nobody but T2 itself touches its exchange-B reader slots.
template <class Tw, class Sync>
__device__ __forceinline__ void fft16_fwd(double (&re)[16], double (&im)[16], cplx *tile, const Tw &twd, int T, Sync sync) {
    radix16_twisted_fwd(re, im);
    twd.template apply16<false>(re, im);
    sync();
    {
        cplx *wp = tile + xa_wbase(T);
#pragma unroll
        for (int p = 0; p < 16; ++p) { cplx v; v.x = re[p]; v.y = im[p]; wp[xa_woff(p)] = v; }
    }
    sync();
    {
        const cplx *rp = tile + xa_rbase(T);
#pragma unroll
        for (int g = 0; g < 16; ++g) { const cplx v = rp[xa_roff(g)]; re[g] = v.x; im[g] = v.y; }
    }
    radix4x4_dif(re, im);
    {
        cplx tw[3];
        twd.load3(tw);
        twiddle4_fwd(re, im, tw);
    }
    __syncwarp();     // from here on everything stays inside this half-warp's region of the tile
    {
        cplx *wp = tile + xb_wbase(T);
#pragma unroll
        for (int g = 0; g < 16; ++g) { cplx v; v.x = re[g]; v.y = im[g]; wp[xb_woff(g)] = v; }
    }
    __syncwarp();
    {
        const cplx *rp = tile + xb_rbase(T);
#pragma unroll
        for (int u = 0; u < 16; ++u) { const cplx v = rp[xb_roff(u)]; re[u] = v.x; im[u] = v.y; }
    }
    radix16_fwd(re, im);
}

// inverse (scaled by 1024): on entry nobody else may be reading this thread's exchange-B reader slots; on exit the tile may
// still be read by the other threads.
template <class Tw, class Sync>
__device__ __forceinline__ void fft16_inv(double (&re)[16], double (&im)[16], cplx *tile, const Tw &twd, int T, Sync sync) {
    radix16_dit_inv(re, im);
    {
        cplx *wp = tile + xb_rbase(T);
#pragma unroll
        for (int u = 0; u < 16; ++u) { cplx v; v.x = re[u]; v.y = im[u]; wp[xb_roff(u)] = v; }
    }
    __syncwarp();
    {
        const cplx *rp = tile + xb_wbase(T);
#pragma unroll
        for (int g = 0; g < 16; ++g) { const cplx v = rp[xb_woff(g)]; re[g] = v.x; im[g] = v.y; }
    }
    {
        cplx tw[3];
        twd.load3(tw);
        twiddle4_inv(re, im, tw);
    }
    radix4x4_dit_inv(re, im);
    __syncwarp();
    {
        cplx *wp = tile + xa_rbase(T);
#pragma unroll
        for (int g = 0; g < 16; ++g) { cplx v; v.x = re[g]; v.y = im[g]; wp[xa_roff(g)] = v; }
    }
    sync();
    {
        const cplx *rp = tile + xa_wbase(T);
#pragma unroll
        for (int p = 0; p < 16; ++p) { const cplx v = rp[xa_woff(p)]; re[p] = v.x; im[p] = v.y; }
    }
    twd.template apply16<true>(re, im);
    radix16_dit_inv(re, im);
    posttwist16_inv(re, im);
}

__device__ __forceinline__ void pack_cplx(double r, double i, uint32_t (&v)[16], int q) {
    v[4 * q] = (uint32_t)__double2loint(r); v[4 * q + 1] = (uint32_t)__double2hiint(r);
    v[4 * q + 2] = (uint32_t)__double2loint(i); v[4 * q + 3] = (uint32_t)__double2hiint(i);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

}  // namespace tb16k
