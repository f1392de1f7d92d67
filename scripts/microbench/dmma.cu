// dmma.cu -- does the FP64 tensor-core instruction (mma.sync.m8n8k4.f64) run on its own pipe next to DFMA on B200?  16 warps per SM,
// each loop iteration issues UN independent DMMAs and / or UF independent DFMAs per thread.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma scripts/microbench/dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4000;

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(512, 1) k(int mode, long long *cycles, double *sink) {
    const int lane = threadIdx.x & 31;
    double c[8][2], f[16];
    for (int i = 0; i < 8; ++i) { c[i][0] = lane * 1e-3 + i; c[i][1] = i; }
    for (int i = 0; i < 16; ++i) f[i] = 1.0 + lane * 1e-3 + i;
    const double a = 1.0 + lane * 1e-6, b = 1.0 - lane * 1e-6;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (mode & 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
        }
        if (mode & 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __fma_rn(f[i], 1.0000001, 1e-9);
        }
    }
    const long long t1 = clock64();
    double acc = 0;
    for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][1];
    for (int i = 0; i < 16; ++i) acc += f[i];
    if (acc == 1.2345) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    long long *cyc; double *sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 8);
    const char *names[] = {"", "DMMA x8", "DFMA x16", "DMMA x8 + DFMA x16"};
    for (int mode = 1; mode <= 3; ++mode) {
        k<<<148, 512>>>(mode, cyc, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += v;
        avg /= 148;
        const double per_iter = avg / ITER;
        const double dmma_flop = (mode & 1) ? 16.0 * 8 * 2 * 8 * 8 * 4 : 0, dfma_flop = (mode & 2) ? 16.0 * 16 * 32 * 2 : 0;   // per SM and iteration
        printf("%-22s %8.1f cycles/iter/SM   DMMA %.1f flop/clk/SM   DFMA %.1f flop/clk/SM\n", names[mode], per_iter, dmma_flop / per_iter, dfma_flop / per_iter);
    }
    return 0;
}
