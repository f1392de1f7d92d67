// probe.cu -- FP64 FMA peak probe: the roofline denominator of the blind-rotation kernels (MEASURED_PEAKS.json carries HBM and bf16
// figures only).  Eight independent dependent-FMA chains per thread, 256 threads per CTA, several CTAs per SM.
#include "kernels.h"

namespace tb {

__global__ void __launch_bounds__(256)
fp64_peak_kernel(double *sink, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) sink[0] = a0;
}

}  // namespace tb

namespace tbk {

cudaError_t launch_fp64_peak(double *sink, int blocks, int iters, cudaStream_t stream) {
    tb::fp64_peak_kernel<<<blocks, 256, 0, stream>>>(sink, iters);
    return cudaGetLastError();
}

}  // namespace tbk
