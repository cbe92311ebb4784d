// diag_kernels.cuh -- measurement helpers (not on the product path).
//
// fma_peak_kernel: the FP32 roofline denominator, measured in the run that quotes it.  Eight independent packed
// accumulator chains per thread of fma.rn.f32x2 (FFMA2: 64 lane-FMAs per warp instruction), 8 CTAs x 256 threads per SM:
// nothing but the FMA pipe is exercised, so lane-FMAs / time is the sustained FP32 issue rate of this GPU at its current
// clocks (148 SMs x 128 lanes x 1.965 GHz = 37.2 T lane-FMA/s nominal).  Same loop as profiles/microbench/ffma2_issue.cu.
#pragma once
#include <cuda_runtime.h>

namespace b200d {

__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, float w, int iters) {
    unsigned long long acc[8], xx, ww;
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    const float x = out[threadIdx.x & 7];
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(xx) : "f"(x));
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %2, %1;" : "+l"(acc[i]) : "l"(xx), "l"(ww));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(acc[i]));
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace b200d
