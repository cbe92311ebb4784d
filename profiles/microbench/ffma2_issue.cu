// Does FFMA2 (fma.rn.f32x2) hold the issue port for one cycle or two?  Mix R independent integer ALU instructions per
// FFMA2 / per pair of scalar FFMAs and time it: if other instructions issue in the shadow of the packed FMA's second
// pipe cycle, R=1 costs nothing extra.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/microbench/ffma2_issue.cu -o /tmp/ffma2_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int R>
__global__ void __launch_bounds__(256) k(float *out, float w, int iters, unsigned seed) {
    unsigned long long acc[8], xx, ww;
    float a[16];
    unsigned z[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = seed + threadIdx.x * 7 + i;
    const float x = out[threadIdx.x & 7];
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(xx) : "f"(x));
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                asm volatile("fma.rn.f32x2 %0, %0, %2, %1;" : "+l"(acc[i]) : "l"(xx), "l"(ww));
            } else {
                asm volatile("fma.rn.f32 %0, %0, %2, %1;" : "+f"(a[2 * i]) : "f"(x), "f"(w));
                asm volatile("fma.rn.f32 %0, %0, %2, %1;" : "+f"(a[2 * i + 1]) : "f"(x), "f"(w));
            }
#pragma unroll
            for (int r = 0; r < R; ++r) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[(i + r) & 7]) : "r"(seed), "r"(it));
        }
    }
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(acc[i]));
    }
    float s = 0;
    unsigned zz = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) zz ^= z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)zz;
}

template <int MODE, int R>
void run(float *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, grid = 148 * 8, block = 256;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<MODE, R><<<grid, block>>>(d, 0.5f, iters, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double fmas = (double)grid * block * iters * 16;
    const double warp_instr = (double)grid * block / 32 * iters * 8 * ((MODE == 0 ? 1 : 2) + R);
    printf("%s + %d LOP3 per 64 lane-FMAs: %.3f ms  %.2f T lane-FMA/s  %.2f warp-instr/cycle/SMSP\n", MODE == 0 ? "FFMA2     " : "2 x FFMA  ", R, ms,
           fmas / ms / 1e9, warp_instr / (ms * 1e-3) / (148 * 4 * 1.965e9));
}

int main() {
    float *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaMemset(d, 0, 148 * 8 * 256 * 4);
    run<0, 0>(d); run<0, 1>(d); run<0, 2>(d); run<0, 3>(d);
    run<1, 0>(d); run<1, 1>(d); run<1, 2>(d);
    return 0;
}
