// Does FFMA2 with an immediate multiplier (what ptxas emits for a constexpr weight) execute on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2_bcast(float &lo, float &hi, float x, float w) {
    unsigned long long acc, xx, ww;
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "f"(lo), "f"(hi));
    asm("mov.b64 %0, {%1,%1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(ww));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
}
template <bool IMM>
__global__ void k(const float *in, float *out, float wrt) {
    float a0 = 0.f, a1 = 1.f;
    const float w = IMM ? (1.0f / 121.0f) : wrt;
    for (int i = 0; i < 8; ++i) ffma2_bcast(a0, a1, in[i + threadIdx.x], w);
    out[2 * threadIdx.x] = a0;
    out[2 * threadIdx.x + 1] = a1;
}
__global__ void kmul(const float *in, float *out) {
    unsigned long long o, a, b;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(in[threadIdx.x]), "f"(in[threadIdx.x + 1]));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(in[threadIdx.x + 2]), "f"(in[threadIdx.x + 3]));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o) : "l"(a), "l"(b));
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(o));
    out[2 * threadIdx.x] = lo; out[2 * threadIdx.x + 1] = hi;
}
int main() {
    float *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, 4096); cudaMemset(in, 0, 4096);
    k<false><<<1, 32>>>(in, out, 1.0f / 121.0f); printf("reg  form: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    kmul<<<1, 32>>>(in, out); printf("mul2 form: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    k<true><<<1, 32>>>(in, out, 0.f); printf("imm  form: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
