// tld4_gather.cu -- can the four bilinear taps of grid_sample(zeros) come from ONE texture-gather instruction?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tld4_gather tld4_gather.cu && ./tld4_gather
// Per pixel and "source": a scattered sampling position (pixel + hashed offset, like a PatchMatch refinement candidate, or
// + a constant shift, like a sweep plane), then
//   A  four predicated LDGs with explicit zero padding (what sample_kernel does today)
//   B  tex2Dgather on a CUDA array (cudaArrayTextureGather), border addressing = the zero padding, coordinates at
//      (floor(ix) + 1, floor(iy) + 1): the footprint is then texels floor(ix), floor(ix)+1 with the hardware's own
//      coordinate at the middle of the two (fraction exactly .5, no fixed-point rounding hazard)
//   C  the same gather on a pitch-linear (cudaResourceTypePitch2D) texture: documented as unsupported -- checked anyway
// Results of B (and C, if it runs) must equal A bit for bit; times by CUDA events.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE, bool SCATTER>
__global__ void __launch_bounds__(256) k(const float *__restrict__ img, cudaTextureObject_t tex, int W, int H, int S, float *out) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    float acc = 0.0f;
    for (int s = 0; s < S; ++s) {
        float ox, oy;
        if (SCATTER) {
            const unsigned h = hash((unsigned)(y * W + x) * 8u + s);
            ox = ((h & 0xffff) * (1.0f / 65536.0f) - 0.5f) * 160.0f;      // +-80 px along the "epipolar line"
            oy = (((h >> 16) & 0xff) * (1.0f / 256.0f) - 0.5f) * 4.0f;
        } else {
            ox = 13.37f * (s + 1) - 40.0f; oy = 0.61f * s - 1.5f;
        }
        const float ix = (float)x + ox, iy = (float)y + oy;
        const float x0 = floorf(ix), y0 = floorf(iy);
        const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
        const float wx1 = ix - x0, wx0 = x1 - ix, wy1 = iy - y0, wy0 = y1 - iy;
        float vnw, vne, vsw, vse;
        if (MODE == 0) {
            const int xi = (int)x0, yi = (int)y0;
            const bool x0in = (unsigned)xi < (unsigned)W, x1in = (unsigned)xi + 1u < (unsigned)W;
            const bool y0in = (unsigned)yi < (unsigned)H, y1in = (unsigned)yi + 1u < (unsigned)H;
            const float *p0 = img + (int)((unsigned)yi * (unsigned)W + (unsigned)xi);
            vnw = (x0in & y0in) ? __ldg(p0) : 0.0f;
            vne = (x1in & y0in) ? __ldg(p0 + 1) : 0.0f;
            vsw = (x0in & y1in) ? __ldg(p0 + W) : 0.0f;
            vse = (x1in & y1in) ? __ldg(p0 + W + 1) : 0.0f;
        } else {
            const float4 g = tex2Dgather<float4>(tex, x1, y1, 0);          // (i,j+1), (i+1,j+1), (i+1,j), (i,j)
            vsw = g.x; vse = g.y; vne = g.z; vnw = g.w;
        }
        float o = __fmul_rn(vnw, __fmul_rn(wx0, wy0));
        o = __fmaf_rn(vne, __fmul_rn(wx1, wy0), o);
        o = __fmaf_rn(vsw, __fmul_rn(wx0, wy1), o);
        o = __fmaf_rn(vse, __fmul_rn(wx1, wy1), o);
        acc += o;
    }
    out[(size_t)y * W + x] = acc;
}

template <int MODE, bool SCATTER>
float run(const float *img, cudaTextureObject_t tex, int W, int H, int S, float *out) {
    dim3 g((W + 31) / 32, (H + 7) / 8), b(32, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k<MODE, SCATTER><<<g, b>>>(img, tex, W, H, S, out);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) k<MODE, SCATTER><<<g, b>>>(img, tex, W, H, S, out);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1.0f; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 20.0f * 1000.0f;
}

int main() {
    const int W = 1600, H = 1200, S = 5;
    std::vector<float> h((size_t)W * H);
    srand(1);
    for (auto &v : h) v = rand() / (float)RAND_MAX;
    float *img, *oA, *oB;
    CK(cudaMalloc(&img, h.size() * 4)); CK(cudaMalloc(&oA, h.size() * 4)); CK(cudaMalloc(&oB, h.size() * 4));
    CK(cudaMemcpy(img, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaArray_t arr;
    CK(cudaMallocArray(&arr, &cd, W, H, cudaArrayTextureGather));
    cudaEvent_t c0, c1; cudaEventCreate(&c0); cudaEventCreate(&c1);
    CK(cudaMemcpy2DToArray(arr, 0, 0, img, W * 4, W * 4, H, cudaMemcpyDeviceToDevice));
    cudaEventRecord(c0);
    for (int i = 0; i < 10; ++i) cudaMemcpy2DToArrayAsync(arr, 0, 0, img, W * 4, W * 4, H, cudaMemcpyDeviceToDevice, 0);
    cudaEventRecord(c1); CK(cudaDeviceSynchronize());
    float cms; cudaEventElapsedTime(&cms, c0, c1);
    printf("copy image -> array: %.1f us per 1600x1200 fp32 image\n", cms * 100.0f);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t texA = 0, texP = 0;
    CK(cudaCreateTextureObject(&texA, &rd, &td, nullptr));
    cudaResourceDesc rp = {}; rp.resType = cudaResourceTypePitch2D; rp.res.pitch2D.devPtr = img; rp.res.pitch2D.desc = cd;
    rp.res.pitch2D.width = W; rp.res.pitch2D.height = H; rp.res.pitch2D.pitchInBytes = (size_t)W * 4;
    cudaError_t ep = cudaCreateTextureObject(&texP, &rp, &td, nullptr);
    printf("pitch2D texture object: %s\n", cudaGetErrorString(ep));
    std::vector<float> a(h.size()), b(h.size());
    for (int sc = 0; sc < 2; ++sc) {
        float tA = sc ? run<0, true>(img, 0, W, H, S, oA) : run<0, false>(img, 0, W, H, S, oA);
        float tB = sc ? run<1, true>(img, texA, W, H, S, oB) : run<1, false>(img, texA, W, H, S, oB);
        CK(cudaMemcpy(a.data(), oA, a.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), oB, b.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0; for (size_t i = 0; i < a.size(); ++i) bad += a[i] != b[i];
        printf("%s: LDG x4 %.1f us   TLD4(array) %.1f us   mismatches %zu of %zu\n", sc ? "scattered" : "regular  ", tA, tB, bad, a.size());
        if (ep == cudaSuccess) {
            float tC = sc ? run<1, true>(img, texP, W, H, S, oB) : run<1, false>(img, texP, W, H, S, oB);
            if (tC >= 0) {
                CK(cudaMemcpy(b.data(), oB, b.size() * 4, cudaMemcpyDeviceToHost));
                bad = 0; for (size_t i = 0; i < a.size(); ++i) bad += a[i] != b[i];
                printf("           TLD4(pitch2D) %.1f us   mismatches %zu\n", tC, bad);
            }
        }
    }
    return 0;
}
