#include <cstdio>
#include <cuda_runtime.h>
// throughput microbenchmark: FFMA (imm), FFMA (reg), FFMA2
template<int MODE>
__global__ void k(float* out, float w, int iters) {
    float a[16];
    #pragma unroll
    for (int i=0;i<16;++i) a[i]=threadIdx.x*0.001f+i;
    float x = out[threadIdx.x & 7];
    if (MODE==0) {
        for (int it=0; it<iters; ++it) {
            #pragma unroll
            for (int i=0;i<16;++i) a[i]=__fmaf_rn(a[i], 0.99826446f, x);
        }
    } else if (MODE==1) {
        for (int it=0; it<iters; ++it) {
            #pragma unroll
            for (int i=0;i<16;++i) a[i]=__fmaf_rn(a[i], w, x);
        }
    } else {
        unsigned long long xx, ww, acc[8];
        asm volatile("mov.b64 %0, {%1,%1};" : "=l"(xx) : "f"(x));
        asm volatile("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
        #pragma unroll
        for (int i=0;i<8;++i) asm volatile("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "f"(a[2*i]), "f"(a[2*i+1]));
        for (int it=0; it<iters; ++it) {
            #pragma unroll
            for (int i=0;i<8;++i) asm volatile("fma.rn.f32x2 %0, %0, %2, %1;" : "+l"(acc[i]) : "l"(xx), "l"(ww));
        }
        #pragma unroll
        for (int i=0;i<8;++i) asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a[2*i]), "=f"(a[2*i+1]) : "l"(acc[i]));
    }
    float s=0; 
    #pragma unroll
    for (int i=0;i<16;++i) s+=a[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
    float* d; cudaMalloc(&d, 148*8*1024*4);
    cudaMemset(d,0,148*8*1024*4);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters=20000; int grid=148*4, block=512;
    for (int mode=0; mode<3; ++mode) {
        for (int rep=0; rep<2; ++rep) {
            cudaEventRecord(e0);
            if (mode==0) k<0><<<grid,block>>>(d,0.5f,iters);
            if (mode==1) k<1><<<grid,block>>>(d,0.5f,iters);
            if (mode==2) k<2><<<grid,block>>>(d,0.5f,iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms,e0,e1);
            double fmas=(double)grid*block*iters*16;
            if (rep) printf("mode %d: %.3f ms  %.2f TFMA/s (lane FMAs)\n", mode, ms, fmas/ms/1e9);
        }
    }
    return 0;
}
