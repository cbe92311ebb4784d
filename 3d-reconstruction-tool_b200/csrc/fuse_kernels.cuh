// fuse_kernels.cuh -- consistency mask + ordered compaction + back-projection to world points.
//
// Reference: PatchMatchMVS._fuse_depth_maps per-view body (src/core/mvs_patchmatch.py:545-565)
// and DenseStereoReconstructor._backproject (src/core/dense_stereo.py:414-437).
// np.where(valid) enumerates pixels in row-major order; the compaction below preserves that
// order (chunk counts -> exclusive scan -> in-chunk scan), so masks, counts and point order are
// bit-identical; coordinates are computed in f64 like the reference's NumPy path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200d {

constexpr int FUSE_NT = 256, FUSE_PPT = 4, FUSE_CHUNK = FUSE_NT * FUSE_PPT;

struct FuseArgs {
    const float *depth, *conf;
    const uint8_t *bgr;
    int H, W;
    float min_conf;
    int mode;                    // 0: conf >= min_conf ; 1: && depth > 0
    double Kinv[9], R[9], t[3];
    double *points;
    uint8_t *colors;
    int64_t *count;
    uint32_t *chunk_counts;
    int64_t *chunk_offsets;
    int n_chunks;
};

__device__ __forceinline__ bool fuse_valid(const FuseArgs &a, size_t i) {
    bool ok = a.conf[i] >= a.min_conf;
    if (a.mode == 1) ok = ok && (a.depth[i] > 0.0f);
    return ok;
}

__global__ void __launch_bounds__(FUSE_NT) fuse_count_kernel(const __grid_constant__ FuseArgs a) {
    const size_t n = (size_t)a.H * a.W;
    const size_t base = (size_t)blockIdx.x * FUSE_CHUNK + (size_t)threadIdx.x * FUSE_PPT;
    int c = 0;
#pragma unroll
    for (int j = 0; j < FUSE_PPT; ++j)
        if (base + j < n && fuse_valid(a, base + j)) ++c;
    __shared__ int s_c[FUSE_NT / 32];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < FUSE_NT / 32; ++w) t += s_c[w];
        a.chunk_counts[blockIdx.x] = (uint32_t)t;
    }
}

// single CTA: exclusive scan of the chunk counts (n_chunks <= a few thousand)
__global__ void __launch_bounds__(1024) fuse_scan_kernel(const __grid_constant__ FuseArgs a) {
    __shared__ int64_t s_part[1024];
    const int tid = threadIdx.x;
    const int per = (a.n_chunks + 1023) / 1024;
    const int b0 = tid * per, b1 = min(b0 + per, a.n_chunks);
    int64_t sum = 0;
    for (int b = b0; b < b1; ++b) sum += a.chunk_counts[b];
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                    // Hillis-Steele inclusive scan
        const int64_t v = (tid >= o) ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int64_t run = s_part[tid] - sum;
    for (int b = b0; b < b1; ++b) {
        a.chunk_offsets[b] = run;
        run += a.chunk_counts[b];
    }
    if (tid == 1023) *a.count = s_part[1023];
}

__global__ void __launch_bounds__(FUSE_NT) fuse_write_kernel(const __grid_constant__ FuseArgs a) {
    const size_t n = (size_t)a.H * a.W;
    const size_t base = (size_t)blockIdx.x * FUSE_CHUNK + (size_t)threadIdx.x * FUSE_PPT;
    bool ok[FUSE_PPT];
    int c = 0;
#pragma unroll
    for (int j = 0; j < FUSE_PPT; ++j) {
        ok[j] = (base + j < n) && fuse_valid(a, base + j);
        c += ok[j];
    }
    // exclusive scan of per-thread counts across the CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __shared__ int s_w[FUSE_NT / 32];
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_w[w];
    int64_t m = a.chunk_offsets[blockIdx.x] + woff + inc - c;
#pragma unroll
    for (int j = 0; j < FUSE_PPT; ++j) {
        if (!ok[j]) continue;
        const size_t i = base + j;
        const int y = (int)(i / a.W), x = (int)(i - (size_t)y * a.W);
        const double d = (double)a.depth[i];
        double p[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double r = (double)x * a.Kinv[k * 3] + (double)y * a.Kinv[k * 3 + 1] + a.Kinv[k * 3 + 2];
            p[k] = r * d - a.t[k];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) a.points[3 * m + k] = p[0] * a.R[0 * 3 + k] + p[1] * a.R[1 * 3 + k] + p[2] * a.R[2 * 3 + k];
        a.colors[3 * m + 0] = a.bgr[3 * i + 2];             // BGR -> RGB
        a.colors[3 * m + 1] = a.bgr[3 * i + 1];
        a.colors[3 * m + 2] = a.bgr[3 * i + 0];
        ++m;
    }
}

}  // namespace b200d
