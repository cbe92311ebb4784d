// sweep_kernels.cuh -- plane-sweep stereo: the whole (plane x neighbour) loop in one kernel.
//
// Reference: DenseStereoReconstructor._plane_sweep_torch / _compute_ncc_torch
// (src/core/dense_stereo.py:262-310, 318-347).  The reference materialises a (D,H,W)
// consistency volume in memory and launches ~45 eager kernels per (plane, neighbour); here a
// CTA keeps its tile of the reference image and its running (best count, best index) on chip
// and sweeps all D planes x N neighbours itself, so HBM sees the images once and three output
// maps -- the path is purely FP32 / shared-memory bound.
//
// Per (plane d, neighbour n): constant-depth reprojection of tile + halo, bilinear sample
// -> smem; window sums (same register-tiled row-major chain as PatchMatch); NCC with the
// plane-sweep epsilon placement  cov / sqrt(var1*var2 + 1e-8)  (:344-345);
// count += (ncc > thresh) & (z > 0.1)  (:303-304).  After the N neighbours of a plane:
// first-max update (torch.max(dim=0) returns the first index of the maximum, :307).
#pragma once
#include "pm_kernels.cuh"

namespace b200d {

struct SweepArgs {
    const float *ref, *mean1, *var1;
    const float *nbr[B200D_MAX_SRC];
    b200d_cams cams;
    int H, W, D;
    const float *depths;                 // (D) fp32, device
    float thresh;
    float *depth_map, *conf;
    int32_t *idx;
    float *ncc;                          // optional (D,N,H,W)
};

template <int K>
__global__ void __launch_bounds__(NT, 2) plane_sweep_kernel(const __grid_constant__ SweepArgs a) {
    using G = Geo<K>;
    constexpr int HALF = G::HALF, RW = G::RW, RH = G::RH, RWP = G::RWP;
    extern __shared__ __align__(16) float smem[];
    float *s_r = smem;
    float *s_s = s_r + RH * RWP;
    float *s_rs = s_s + RH * RWP;
    unsigned char *s_valid = reinterpret_cast<unsigned char *>(s_rs + RH * RWP);

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = a.H, W = a.W, N = a.cams.n_src;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);

    for (int i = tid; i < RH * RW; i += NT) {
        const int ry = i / RW, rx = i - ry * RW;
        const int gx = x0 - HALF + rx, gy = y0 - HALF + ry;
        const bool in = (gx >= 0) && (gx < W) && (gy >= 0) && (gy < H);
        s_r[ry * RWP + rx] = in ? __ldg(a.ref + (size_t)gy * W + gx) : 0.0f;
    }
    if constexpr (RWP > RW) {
        for (int i = tid; i < RH * (RWP - RW); i += NT) {
            const int ry = i / (RWP - RW), rx = RW + i % (RWP - RW);
            s_s[ry * RWP + rx] = 0.0f;
            s_rs[ry * RWP + rx] = 0.0f;
        }
    }

    float m1[QY][PX], v1[QY][PX], bestc[QY][PX];
    int besti[QY][PX];
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int gx = x0 + tx * PX + p, gy = y0 + ty * QY + q;
            const bool in = gx < W && gy < H;
            m1[q][p] = in ? __ldg(a.mean1 + (size_t)gy * W + gx) : 0.0f;
            v1[q][p] = in ? __ldg(a.var1 + (size_t)gy * W + gx) : 0.0f;
            bestc[q][p] = -1.0f;
            besti[q][p] = 0;
        }
    __syncthreads();

    float as[QY][PX], ass[QY][PX], ars[QY][PX], count[QY][PX];

    for (int d = 0; d < a.D; ++d) {
        const float depth = __ldg(a.depths + d);
#pragma unroll
        for (int q = 0; q < QY; ++q)
#pragma unroll
            for (int p = 0; p < PX; ++p) count[q][p] = 0.0f;

        for (int n = 0; n < N; ++n) {
            const float *__restrict__ img = a.nbr[n];
            for (int i = tid; i < RH * RW; i += NT) {
                const int ry = i / RW, rx = i - ry * RW;
                const int gx = x0 - HALF + rx, gy = y0 - HALF + ry;
                float sv = 0.0f, rsv = 0.0f;
                if ((gx >= 0) && (gx < W) && (gy >= 0) && (gy < H)) {
                    float pw[3], u, v, z;
                    backproject_world(a.cams, (float)gx, (float)gy, depth, pw);
                    project_src(a.cams, n, pw, u, v, z);
                    sv = sample_bilinear(img, W, wm1, hm1, u, v);
                    rsv = mul(s_r[ry * RWP + rx], sv);
                    const int ox = rx - HALF, oy = ry - HALF;
                    if (ox >= 0 && ox < TW && oy >= 0 && oy < TH) s_valid[oy * TW + ox] = (z > 0.1f);
                }
                s_s[ry * RWP + rx] = sv;
                s_rs[ry * RWP + rx] = rsv;
            }
            __syncthreads();
            box_sums<K>(s_s, s_rs, tx, ty, as, ass, ars);
#pragma unroll
            for (int q = 0; q < QY; ++q)
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const int ox = tx * PX + p, oy = ty * QY + q;
                    const int gx = x0 + ox, gy = y0 + oy;
                    if (gx < W && gy < H) {
                        const float m2 = as[q][p];
                        const float v2 = sub(ass[q][p], mul(m2, m2));
                        const float cov = sub(ars[q][p], mul(m1[q][p], m2));
                        const float ncc = div_(cov, __fsqrt_rn(add(mul(v1[q][p], v2), 1e-8f)));   // :344-345
                        if (a.ncc) a.ncc[(((size_t)d * N + n) * H + gy) * W + gx] = ncc;
                        if ((ncc > a.thresh) && s_valid[oy * TW + ox]) count[q][p] = add(count[q][p], 1.0f);
                    }
                }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < QY; ++q)
#pragma unroll
            for (int p = 0; p < PX; ++p)
                if (count[q][p] > bestc[q][p]) { bestc[q][p] = count[q][p]; besti[q][p] = d; }
    }

#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int gx = x0 + tx * PX + p, gy = y0 + ty * QY + q;
            if (gx < W && gy < H) {
                const size_t g = (size_t)gy * W + gx;
                a.depth_map[g] = __ldg(a.depths + besti[q][p]);      // depths_tensor[best_depth_idx]  (:310)
                a.conf[g] = bestc[q][p];
                a.idx[g] = besti[q][p];
            }
        }
}

}  // namespace b200d
