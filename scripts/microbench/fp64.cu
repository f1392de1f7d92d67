// fp64.cu -- FP64 issue behaviour on B200: cycles per DFMA as a function of warps per SM and independent chains per thread.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64 scripts/microbench/fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4000;
template <int CHAINS, int KIND>
__global__ void k(long long *cycles, double *sink) {
    double f[CHAINS];
    for (int c = 0; c < CHAINS; ++c) f[c] = 1.0 + threadIdx.x * 1e-3 + c;
    const double m = 1.0000001 + blockIdx.x * 1e-12, a = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (KIND == 0) f[c] = __fma_rn(f[c], m, a);
            if (KIND == 1) f[c] = __dadd_rn(f[c], a);
            if (KIND == 2) f[c] = __dmul_rn(f[c], m);
        }
    }
    const long long t1 = clock64();
    double acc = 0;
    for (int c = 0; c < CHAINS; ++c) acc += f[c];
    if (acc == 1.2345) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int CHAINS, int KIND>
void run(const char *name, int warps, long long *cyc, double *sink) {
    k<CHAINS, KIND><<<148, warps * 32>>>(cyc, sink);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += v;
    avg /= 148;
    const double per_iter = avg / ITER;
    printf("%s chains=%2d warps/SM=%2d: %8.1f cycles per iteration per warp, %6.2f cycles per instr per warp, %5.2f warp-instr/clk/SM (peak 2.0)\n",
           name, CHAINS, warps, per_iter, per_iter / CHAINS, warps * CHAINS / per_iter);
}
int main() {
    long long *cyc; double *sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 8);
    for (int warps : {1, 4, 8, 16, 32}) {
        run<1, 0>("DFMA", warps, cyc, sink);
        run<4, 0>("DFMA", warps, cyc, sink);
        run<16, 0>("DFMA", warps, cyc, sink);
        run<32, 0>("DFMA", warps, cyc, sink);
    }
    for (int warps : {4, 16}) { run<16, 1>("DADD", warps, cyc, sink); run<16, 2>("DMUL", warps, cyc, sink); }
    return 0;
}
