// f32x2_contract.cu -- ptxas 12.9 contracts `mul.rn.f32x2` + `add.rn.f32x2` into one FFMA2 (the scalar forms stay apart).
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -cubin -o f.cubin f32x2_contract.cu && cuobjdump -sass f.cubin | grep -E "FMUL|FADD|FFMA"
// Output here (CUDA 12.9.86):   kernel k  (packed):  FFMA2 R8, R8.F32x2.HI_LO, R10.F32x2.HI_LO, R12.F32x2.HI_LO      <- contracted
//                               kernel ks (scalar):  FMUL, FMUL, FADD, FADD                                           <- as written
// Explicit .rn and -fmad=false (nvcc forwards `--fmad false` to ptxas) do not prevent it; rewriting the product as
// fma(a, b, -0.0) does not either (ptxas folds it back to a multiply first).  Consequence for the parity kernels: a packed
// product must never be the direct operand of a packed add / sub (device_math.cuh, "TOOLCHAIN CAVEAT").
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b){unsigned long long r; asm("mul.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b){unsigned long long r; asm("add.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b)); return r;}
__global__ void k(const float2* in, float2* out){
  int i = blockIdx.x*blockDim.x+threadIdx.x;
  float2 a=in[3*i], b=in[3*i+1], c=in[3*i+2];
  unsigned long long r = add2(mul2(pk(a.x,a.y), pk(b.x,b.y)), pk(c.x,c.y));
  float lo, hi; asm("mov.b64 {%0,%1}, %2;":"=f"(lo),"=f"(hi):"l"(r));
  out[i] = make_float2(lo,hi);
}
__global__ void ks(const float2* in, float2* out){
  int i = blockIdx.x*blockDim.x+threadIdx.x;
  float2 a=in[3*i], b=in[3*i+1], c=in[3*i+2];
  out[i] = make_float2(__fadd_rn(__fmul_rn(a.x,b.x),c.x), __fadd_rn(__fmul_rn(a.y,b.y),c.y));
}
