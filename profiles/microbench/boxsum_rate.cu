// Microbenchmark: the window-sum inner loop of ncc_kernel (box_sums<K>) alone, fed from shared memory that is filled
// once -- no TMA, no epilogue, no tile transitions.  Reports lane-FMA/s of the chains (3*K*K per output) so the loss
// between this and the full kernel can be attributed.
//   nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a -I include -I 3d-reconstruction-tool_b200/csrc \
//        profiles/microbench/boxsum_rate.cu -o /tmp/boxsum_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ncc_kernels.cuh"
using namespace b200d;

template <int K, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_box(float *out, int iters) {
    using G = Geo<K>;
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int BUF = G::buf_bytes / sizeof(float);
    float *s_r = reinterpret_cast<float *>(sm);
    for (int i = threadIdx.x; i < 3 * BUF; i += NT) s_r[i] = (float)((i * 2654435761u) >> 8) * (1.0f / 16777216.0f);
    __syncthreads();
    const int tx = threadIdx.x % NTX, ty = threadIdx.x / NTX;
    float as[QY][PX], ass[QY][PX], ars[QY][PX], tot = 0.f;
    for (int it = 0; it < iters; ++it) {
        box_sums<K>(s_r + (1 + (it & 1)) * BUF, s_r, tx, ty, as, ass, ars);
#pragma unroll
        for (int q = 0; q < QY; ++q)
#pragma unroll
            for (int p = 0; p < PX; ++p) tot = add(tot, add(as[q][p], add(ass[q][p], ars[q][p])));
    }
    out[blockIdx.x * NT + threadIdx.x] = tot;
}

template <int K, int MINB>
void run(int ctas_per_sm, int iters) {
    float *d;
    cudaMalloc(&d, 148 * 8 * NT * 4);
    cudaFuncSetAttribute(k_box<K, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<K>::smem_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_box<K, MINB><<<grid, NT, Geo<K>::smem_bytes>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)grid * NT * PX * QY * 3.0 * K * K * iters;
        if (rep) printf("K=%d ctas/SM=%d (launch bound %d): %.3f ms  %.2f T lane-FMA/s (chains only)  err=%s\n", K, ctas_per_sm, MINB, ms,
                        fma / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d);
}

int main() {
    run<11, 3>(3, 400);
    run<11, 3>(2, 400);
    run<11, 3>(1, 400);
    run<11, 2>(2, 400);
    run<7, 3>(3, 800);
    run<5, 3>(3, 1200);
    return 0;
}
