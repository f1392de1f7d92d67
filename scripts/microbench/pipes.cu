// pipes.cu -- per-SM throughput of the pipes the blind-rotation kernels compete for (B200, sm_100a): SHFL, LDS.128, STS.128, LDTM/STTM,
// DFMA, and pairs of them running concurrently.  One CTA of 512 threads (16 warps) per SM, 148 CTAs; each test runs ITER iterations
// of UNROLL independent operations per thread and reports SM cycles per warp-instruction (averaged over the SM's 16 warps).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipes scripts/microbench/pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 2000;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ double2 lds128(const void *p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ double lds64(const void *p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(void *p, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(smem_u32(p)), "d"(v.x), "d"(v.y) : "memory"); }

// mode bits: 1 = SHFL x8, 2 = LDS.128 x8, 4 = STS.128 x8, 8 = LDTM.x16 x2, 16 = DFMA x16, 32 = STTM.x16 x2, 64 = SEL x16 (ALU), 128 = LDS.64 x8
__global__ void __launch_bounds__(512, 1) pipes(int mode, long long *cycles, double *sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t tmem_base;
    const int W = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (W == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)((W & 3) * 32) << 16) + (uint32_t)((W >> 2) * 64);
    double2 *sp = reinterpret_cast<double2 *>(smem) + threadIdx.x;      // conflict-free 16-byte accesses, 512 threads x 8 slots = 64 KiB
    uint32_t s[8] = {1u * lane, 2u + lane, 3u, 4u, 5u, 6u, 7u, 8u};
    double2 l[8];
    for (int k = 0; k < 8; ++k) { l[k] = make_double2(lane + k, k); sp[k * 512] = l[k]; }
    uint32_t tv[16];
    for (int k = 0; k < 16; ++k) tv[k] = lane + k;
    tmem_st16(taddr, tv); tmem_st16(taddr + 16, tv);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    double f[16];
    for (int k = 0; k < 16; ++k) f[k] = 1.0 + lane * 1e-3 + k;
    uint32_t sel[16];
    for (int k = 0; k < 16; ++k) sel[k] = lane * 7 + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (mode & 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] = __shfl_xor_sync(0xffffffffu, s[k], 16);
        }
        if (mode & 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { double2 v = lds128(sp + k * 512); l[k].x = v.x; l[k].y = v.y; }
        }
        if (mode & 128) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { l[k].x = lds64(reinterpret_cast<double *>(smem) + threadIdx.x + k * 512); }
        }
        if (mode & 4) {
#pragma unroll
            for (int k = 0; k < 8; ++k) sts128(sp + k * 512, l[k]);
        }
        if (mode & 8) {
            uint32_t a[16], b[16];
            tmem_ld16(taddr, a); tmem_ld16(taddr + 16, b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 16; ++k) tv[k] ^= a[k] + b[k];
        }
        if (mode & 32) {
            tmem_st16(taddr, tv); tmem_st16(taddr + 16, tv);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (mode & 16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = __fma_rn(f[k], 1.0000001, 1e-9);
        }
        if (mode & 64) {
#pragma unroll
            for (int k = 0; k < 16; ++k) sel[k] = (sel[k] & 1) ? sel[(k + 1) & 15] : sel[(k + 5) & 15] + 1;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    double acc = 0;
    for (int k = 0; k < 8; ++k) acc += s[k] + l[k].x + l[k].y;
    for (int k = 0; k < 16; ++k) acc += f[k] + tv[k] + sel[k];
    if (acc == 1.2345) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (W == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

int main() {
    long long *cyc; double *sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(pipes, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    struct { int mode; const char *name; int warp_instr; } tests[] = {
        {1, "SHFL x8", 8}, {2, "LDS.128 x8", 8}, {128, "LDS.64 x8", 8}, {4, "STS.128 x8", 8}, {8, "LDTM.x16 x2", 2}, {32, "STTM.x16 x2", 2}, {16, "DFMA x16", 16},
        {64, "SEL-ish ALU x16", 16}, {1 | 2, "SHFL x8 + LDS.128 x8", 16}, {2 | 4, "LDS.128 x8 + STS.128 x8", 16}, {2 | 8, "LDS.128 x8 + LDTM.x16 x2", 10},
        {2 | 16, "LDS.128 x8 + DFMA x16", 24}, {1 | 16, "SHFL x8 + DFMA x16", 24}, {8 | 16, "LDTM.x16 x2 + DFMA x16", 18},
        {2 | 4 | 8 | 16, "LDS x8 + STS x8 + LDTM x2 + DFMA x16", 34}, {16 | 64, "DFMA x16 + ALU x16", 32}, {1 | 4, "SHFL x8 + STS.128 x8", 16},
    };
    for (auto &t : tests) {
        pipes<<<148, 512, 65536>>>(t.mode, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", t.name, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += v;
        avg /= 148;
        // SM cycles per iteration, and per warp-instruction summed over the 16 warps of the SM
        printf("%-44s %9.1f cycles/iter/SM  = %6.2f cycles per warp-instruction (16 warps x %d instr per iter)\n", t.name, avg / ITER, avg / ITER / (16.0 * t.warp_instr), t.warp_instr);
    }
    return 0;
}
