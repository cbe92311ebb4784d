// Minimal TMA 2D tile load with negative coordinates / OOB zero fill, descriptor nested in a __grid_constant__ struct.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
struct Args { CUtensorMap tm[2]; int W, H; float *out; int bw, bh, cx, cy; };
__device__ __forceinline__ unsigned su(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char sm[];
    float *buf = (float *)sm;
    unsigned long long *bar = (unsigned long long *)(sm + 32768);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su(bar)), "r"(a.bw * a.bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su(buf)), "l"(&a.tm[1]), "r"(su(bar)), "r"(a.cx), "r"(a.cy) : "memory");
    }
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(su(bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < a.bw * a.bh; i += blockDim.x) a.out[i] = buf[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int W = 64, H = 48, bw = 68, bh = 34;
    std::vector<float> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = 1.0f + i;
    float *d, *o; cudaMalloc(&d, W * H * 4); cudaMalloc(&o, bw * bh * 4);
    cudaMemcpy(d, h.data(), W * H * 4, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Enc enc = (Enc)p;
    Args a; memset(&a, 0, sizeof(a));
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * 4};
    cuuint32_t box[2] = {bw, bh}, es[2] = {1, 1};
    CUresult r = enc(&a.tm[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d (q=%d)\n", (int)r, (int)q);
    a.W = W; a.H = H; a.out = o; a.bw = bw; a.bh = bh;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    int tests[1][2] = {{argc > 1 ? atoi(argv[1]) : 0, argc > 2 ? atoi(argv[2]) : 0}};
    for (auto &t : tests) {
        a.cx = t[0]; a.cy = t[1];
        k<<<1, 128, 40000>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> ho(bw * bh);
        cudaMemcpy(ho.data(), o, bw * bh * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
            int gx = x + t[0], gy = y + t[1];
            float exp = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[gy * W + gx] : 0.f;
            bad += ho[y * bw + x] != exp;
        }
        printf("coords (%d,%d): %s, mismatches %d\n", t[0], t[1], cudaGetErrorString(e), bad);
    }
    return 0;
}
