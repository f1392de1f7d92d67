// tmem_cp.cu -- can tcgen05.cp (shared memory -> Tensor Memory, asynchronous proxy) carry the read half of an FFT exchange?
//   1. layout: element r (16 bytes) of a contiguous 2 KiB block lands in TMEM lane r (128x128b, no swizzle, SBO = 128 bytes);
//   2. throughput of back-to-back copies issued by one thread (2 KiB and 4 KiB shapes);
//   3. the same while the other 15 warps stream LDS.128 / STS.128: does the copy engine share the LSU's shared-memory bandwidth?
//   4. latency of one whole exchange step: STS.128 x16 -> fence.proxy.async -> barrier -> 16 copies -> commit -> mbarrier -> LDTM x4.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_cp scripts/microbench/tmem_cp.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ double2 lds128(const void *p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(void *p, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(smem_u32(p)), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void cp_128x128b(uint32_t taddr, uint64_t desc) { asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory"); }
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t desc) { asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory"); }
__device__ __forceinline__ void commit(void *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try(void *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) if (clock64() - t0 > 2000000000LL) __trap();
}

constexpr int ITER = 500;

// mode 0: layout check; 1: 16 x cp 128x128b per commit, alone; 2: 8 x cp 128x256b per commit, alone; 3: mode 1 + 15 warps of LDS.128;
// 4: mode 1 + 15 warps of STS.128; 5: 15 warps of LDS.128 alone (baseline for 3); 6: 15 warps STS.128 alone; 7: whole exchange step (all 16 warps in 4 groups)
__global__ void __launch_bounds__(512, 1) k(int mode, long long *cycles, uint32_t *check) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base;
    __shared__ unsigned long long bar[4];
    const int W = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (W == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    double2 *sp = reinterpret_cast<double2 *>(smem);
    for (int i = threadIdx.x; i < 4096; i += 512) sp[i] = make_double2((double)i, (double)(i + 0.5));   // 64 KiB
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    const uint32_t quarter = tb + ((uint32_t)((W & 3) * 32) << 16);
    double2 acc = make_double2(0, 0);
    long long t0 = clock64(), t1 = t0;

    if (mode == 0) {
        if (threadIdx.x == 0) {
            cp_128x128b(tb + 0, desc_noswz(smem_u32(sp), 128, 128));                 // elements 0..127 -> lanes 0..127, columns 0-3
            cp_128x256b(tb + 8, desc_noswz(smem_u32(sp + 128), 2048, 128));          // 32 bytes per lane: K chunk 0 at +0, chunk 1 at +LBO
            commit(&bar[0]);
        }
        mbar_wait(&bar[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (W < 4) {
            uint32_t v[16];
            tmem_ld16(quarter, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 16; ++c) check[(W * 32 + lane) * 16 + c] = v[c];
        }
    } else if (mode >= 1 && mode <= 6) {
        const bool copier = (mode <= 4) && threadIdx.x == 0;
        const bool lds = (mode == 3 || mode == 5) && W >= 1, sts = (mode == 4 || mode == 6) && W >= 1;
        __syncthreads();
        t0 = clock64();
        if (copier) {
            uint32_t ph = 0;
            for (int it = 0; it < ITER; ++it) {
                if (mode == 2) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) cp_128x256b(tb + 8 * c, desc_noswz(smem_u32(sp + 256 * c), 2048, 128));
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) cp_128x128b(tb + 4 * c, desc_noswz(smem_u32(sp + 128 * c), 128, 128));
                }
                commit(&bar[0]);
                mbar_wait(&bar[0], ph);
                ph ^= 1u;
            }
        } else if (lds) {
            const double2 *p = sp + (threadIdx.x - 32);
            for (int it = 0; it < ITER * 4; ++it) {
#pragma unroll
                for (int q = 0; q < 8; ++q) { double2 v = lds128(p + q * 480); acc.x += v.x; acc.y += v.y; }
            }
        } else if (sts) {
            double2 *p = sp + 2048 + (threadIdx.x - 32);   // upper 32 KiB: not the region being copied
            for (int it = 0; it < ITER * 4; ++it) {
#pragma unroll
                for (int q = 0; q < 4; ++q) sts128(p + q * 480, acc);
                acc.x += 1.0;
            }
        }
        t1 = clock64();
    } else if (mode == 7) {
        // four groups of four warps (one "ciphertext" each: 128 threads = 128 TMEM lanes), staggered; per step: 16 STS.128 per thread, 16 copies, 4 LDTM.x16
        const int grp = W >> 2;
        double2 *tile = sp + grp * 1024 * 0;   // all groups write the same 32 KiB region pattern? no: 16 KiB per group
        tile = sp + grp * 1024;
        uint32_t ph = 0;
        const uint32_t land = tb + (uint32_t)(grp * 64);
        __syncthreads();
        t0 = clock64();
        for (int it = 0; it < ITER; ++it) {
#pragma unroll
            for (int m = 0; m < 8; ++m) sts128(tile + m * 128 + (threadIdx.x & 127), acc);    // half an exchange: 8 registers x 128 threads = 16 KiB
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            if ((threadIdx.x & 127) == 0) {
#pragma unroll
                for (int m = 0; m < 8; ++m) cp_128x128b(land + 4 * m, desc_noswz(smem_u32(tile + m * 128), 128, 128));
                commit(&bar[grp]);
            }
            mbar_wait(&bar[grp], ph);
            ph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t a[16], b[16];
            tmem_ld16(quarter + (uint32_t)(grp * 64), a);
            tmem_ld16(quarter + (uint32_t)(grp * 64) + 16, b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc.x += __hiloint2double(a[1], a[0]) + __hiloint2double(b[3], b[2]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");    // everyone has read: the landing columns and the tile may be rewritten
        }
        t1 = clock64();
    } else if (mode == 8) {
        // the same step through shared memory only: 8 STS.128, barrier, 8 LDS.128, barrier
        const int grp = W >> 2;
        double2 *tile = sp + grp * 1024;
        __syncthreads();
        t0 = clock64();
        for (int it = 0; it < ITER; ++it) {
#pragma unroll
            for (int m = 0; m < 8; ++m) sts128(tile + m * 128 + (threadIdx.x & 127), acc);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
            for (int m = 0; m < 8; ++m) { double2 v = lds128(tile + m * 128 + ((threadIdx.x + 32) & 127)); acc.x += v.x; acc.y += v.y; }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        }
        t1 = clock64();
    }
    __syncthreads();
    if (acc.x == 1.2345) check[0] = 1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (threadIdx.x == 32) cycles[148 + blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (W == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

int main() {
    long long *cyc; uint32_t *chk;
    cudaMalloc(&cyc, 2 * 148 * 8); cudaMalloc(&chk, 128 * 16 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    auto run = [&](int mode, int grid, double *c0, double *c1) {
        k<<<grid, 512, 65536 + 1024>>>(mode, cyc, chk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); exit(1); }
        long long h[296];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double a = 0, b = 0;
        for (int i = 0; i < grid; ++i) { a += h[i]; b += h[148 + i]; }
        *c0 = a / grid; *c1 = b / grid;
    };
    double c0, c1;
    run(0, 1, &c0, &c1);
    {
        static uint32_t h[128 * 16];
        cudaMemcpy(h, chk, sizeof h, cudaMemcpyDeviceToHost);
        int bad128 = 0, bad256 = 0;
        for (int r = 0; r < 128; ++r) {
            double x, y;
            uint64_t wx = ((uint64_t)h[r * 16 + 1] << 32) | h[r * 16], wy = ((uint64_t)h[r * 16 + 3] << 32) | h[r * 16 + 2];
            memcpy(&x, &wx, 8); memcpy(&y, &wy, 8);
            if (x != (double)r || y != r + 0.5) ++bad128;
            // 128x256b: columns 8..15 of lane r = 32 bytes: chunk 0 = element 128 + r (at +0), chunk 1 = element at +LBO = 128 + 128 + r
            uint64_t w0 = ((uint64_t)h[r * 16 + 9] << 32) | h[r * 16 + 8], w2 = ((uint64_t)h[r * 16 + 13] << 32) | h[r * 16 + 12];
            double e0, e2; memcpy(&e0, &w0, 8); memcpy(&e2, &w2, 8);
            if (e0 != (double)(128 + r) || e2 != (double)(256 + r)) ++bad256;
            if (r < 3 || r == 127) printf("  lane %3d: 128b -> (%g, %g)   256b -> (%g ..., %g ...)\n", r, x, y, e0, e2);
        }
        printf("layout 128x128b (SBO 128): %d of 128 lanes wrong;  128x256b (LBO 2048, SBO 128): %d wrong\n", bad128, bad256);
    }
    run(1, 148, &c0, &c1); printf("16 x cp.128x128b + commit + wait, alone:        %8.1f cycles per 32 KiB  = %.1f B/clk\n", c0 / ITER, 32768.0 / (c0 / ITER));
    run(2, 148, &c0, &c1); printf(" 8 x cp.128x256b + commit + wait, alone:        %8.1f cycles per 32 KiB  = %.1f B/clk\n", c0 / ITER, 32768.0 / (c0 / ITER));
    double l_alone, s_alone;
    run(5, 148, &c0, &l_alone); printf("15 warps LDS.128 alone:                          %8.1f cycles per 8 LDS.128 x 15 warps = %.1f B/clk\n", l_alone / (ITER * 4), 15 * 8 * 512.0 / (l_alone / (ITER * 4)));
    run(3, 148, &c0, &c1); printf("copies + 15 warps LDS.128:   copies %8.1f cycles per 32 KiB (%.1f B/clk);  LDS %.1f B/clk\n", c0 / ITER, 32768.0 / (c0 / ITER), 15 * 8 * 512.0 / (c1 / (ITER * 4)));
    run(6, 148, &c0, &s_alone); printf("15 warps STS.128 alone:                          %8.1f cycles per 4 STS.128 x 15 warps = %.1f B/clk\n", s_alone / (ITER * 4), 15 * 4 * 512.0 / (s_alone / (ITER * 4)));
    run(4, 148, &c0, &c1); printf("copies + 15 warps STS.128:   copies %8.1f cycles per 32 KiB (%.1f B/clk);  STS %.1f B/clk\n", c0 / ITER, 32768.0 / (c0 / ITER), 15 * 4 * 512.0 / (c1 / (ITER * 4)));
    run(7, 148, &c0, &c1); printf("half-exchange via cp (8 STS, fence, bar, 8 cp, commit, wait, 2 LDTM.x16, bar), 4 groups: %8.1f cycles per step\n", c0 / ITER);
    run(8, 148, &c0, &c1); printf("half-exchange via smem (8 STS, bar, 8 LDS, bar), 4 groups:                               %8.1f cycles per step\n", c0 / ITER);
    return 0;
}
