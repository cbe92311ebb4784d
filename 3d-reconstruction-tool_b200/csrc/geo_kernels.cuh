// geo_kernels.cuh -- cross-view kernels (EXTENSIONS: the reference has no counterpart, so no dackey-wav oracle).
//
// The reference's "multi-view consistency" is photometric (`_compute_confidence` counts sources with
// NCC > 0.6) and its README's "view propagation" has no code.  BASELINE.json's north_star asks for both
// in their usual form, fed by the depth / normal maps of the other views (all-gathered over NCCL when views are
// sharded).  Both kernels share one forward-backward round trip, written with the camera helpers of
// device_math.cuh (same fp32 operation sequence as the sampling path: K with skew, fp32 Kinv, R/t per view), so
// the definition restated in oracle/dense_oracle.c (orc_geo_consistency / orc_viewprop_candidates) is reproduced
// bit for bit (tests/test_gpu_geo.py, tests/test_gpu_viewprop.py):
//
//   X   = backproject(ref; p, d)                             world point of reference pixel p at depth d
//   q   = X in source camera;  needs q.z > 1e-6
//   p_s = rint(K (q.xy / (q.z + 1e-8)))                      nearest source pixel, must lie inside the image
//   d_s = depth_src[p_s];  needs d_s > 0
//   X'  = backproject(src; p_s, d_s);   r = X' in reference camera;  needs r.z > 1e-6
//   p'  = K (r.xy / (r.z + 1e-8))
//
// geo_consistency_kernel   source s is consistent  <=>  |p' - p|^2 < px_thresh^2  and  |r.z - d| < rel_thresh * d ;
//                          out[p] = number of consistent sources.
// viewprop_cand_kernel     candidate hypothesis of p from source s: depth r.z (if depth_min <= r.z <= depth_max) and
//                          normal R_ref R_s^T n_s[p_s]; where the round trip fails the pixel keeps its hypothesis and
//                          may not adopt (allow = 0).
// One thread per pixel, coalesced; HBM / L2-bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_math.cuh"

namespace b200d {

struct GeoArgs {
    const float *depth_ref;
    const float *depth_src[B200D_MAX_SRC];
    b200d_cams cams;
    int H, W;
    float px_thresh, rel_thresh;
    float *out;
};

struct RoundTrip {
    bool ok;
    int us, vs;          // nearest source pixel
    float rz;            // depth of the source's surface point in the reference camera
    float du, dv;        // reprojection error in the reference image
};

// forward-backward reprojection of reference pixel (x,y) at depth d through source s and its depth map
__device__ __forceinline__ RoundTrip round_trip(const b200d_cams &c, int s, const float *__restrict__ depth_src, int W,
                                                int H, int x, int y, float d, const float Xw[3]) {
    RoundTrip r;
    r.ok = false;
    r.us = r.vs = 0;
    r.rz = r.du = r.dv = 0.0f;
    float q[3];
    world_to_cam(c.R_src[s], c.t_src[s], Xw, q);
    if (!(q[2] > 1e-6f)) return r;
    const float zz = add(q[2], 1e-8f);
    float u, v;
    to_pixel(c, div_(q[0], zz), div_(q[1], zz), u, v);
    const float uf = rintf(u), vf = rintf(v);                        // round half to even, like np.rint
    if (!((uf >= 0.0f) & (uf <= (float)(W - 1)) & (vf >= 0.0f) & (vf <= (float)(H - 1)))) return r;   // NaN fails
    r.us = (int)uf;
    r.vs = (int)vf;
    const float ds = __ldg(depth_src + (size_t)r.vs * W + r.us);
    if (!(ds > 0.0f)) return r;
    float Xb[3], pr[3];
    backproject_cam(c.Kinv, c.R_src[s], c.t_src[s], uf, vf, ds, Xb);
    world_to_cam(c.R_ref, c.t_ref, Xb, pr);
    if (!(pr[2] > 1e-6f)) return r;
    const float zr = add(pr[2], 1e-8f);
    float pu, pv;
    to_pixel(c, div_(pr[0], zr), div_(pr[1], zr), pu, pv);
    r.rz = pr[2];
    r.du = sub(pu, (float)x);
    r.dv = sub(pv, (float)y);
    r.ok = true;
    return r;
}

__global__ void __launch_bounds__(256) geo_consistency_kernel(const __grid_constant__ GeoArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const b200d_cams &c = a.cams;
    const size_t g = (size_t)y * a.W + x;
    const float d = a.depth_ref[g];
    float count = 0.0f;
    if (d > 0.0f) {
        float Xw[3];
        backproject_cam(c.Kinv, c.R_ref, c.t_ref, (float)x, (float)y, d, Xw);
        const float px2 = mul(a.px_thresh, a.px_thresh), dtol = mul(a.rel_thresh, d);
        for (int s = 0; s < c.n_src; ++s) {
            const RoundTrip r = round_trip(c, s, a.depth_src[s], a.W, a.H, x, y, d, Xw);
            const bool ok = r.ok && (add(mul(r.du, r.du), mul(r.dv, r.dv)) < px2) && (fabsf(sub(r.rz, d)) < dtol);
            count = add(count, ok ? 1.0f : 0.0f);
        }
    }
    a.out[g] = count;
}

struct VpropArgs {
    const float *depth, *normal;             // current hypothesis of the reference view (H,W), (H,W,3)
    const float *depth_src, *normal_src;     // snapshot of source s's maps
    b200d_cams cams;
    int s, H, W;
    float depth_min, depth_max;
    float *cand_depth, *cand_normal;         // out (H,W), (H,W,3)
    uint8_t *allow;                          // out (H,W): 1 = the candidate is a real proposal and may be adopted
};

__global__ void __launch_bounds__(256) viewprop_cand_kernel(const __grid_constant__ VpropArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const b200d_cams &c = a.cams;
    const size_t g = (size_t)y * a.W + x;
    const float d = a.depth[g];
    float cd = d, n0 = a.normal[3 * g], n1 = a.normal[3 * g + 1], n2 = a.normal[3 * g + 2];
    uint8_t allow = 0;
    if (d > 0.0f) {
        float Xw[3];
        backproject_cam(c.Kinv, c.R_ref, c.t_ref, (float)x, (float)y, d, Xw);
        const RoundTrip r = round_trip(c, a.s, a.depth_src, a.W, a.H, x, y, d, Xw);
        if (r.ok && (r.rz >= a.depth_min) && (r.rz <= a.depth_max)) {
            const float *ns = a.normal_src + 3 * ((size_t)r.vs * a.W + r.us);
            const float s0 = __ldg(ns), s1 = __ldg(ns + 1), s2 = __ldg(ns + 2);
            const float *Rs = c.R_src[a.s];
            float nw[3];                                             // n_w = R_s^T n_s
#pragma unroll
            for (int n = 0; n < 3; ++n) nw[n] = fma_(s2, Rs[2 * 3 + n], fma_(s1, Rs[1 * 3 + n], mul(s0, Rs[0 * 3 + n])));
            // n_r = R_ref n_w
            n0 = fma_(nw[2], c.R_ref[2], fma_(nw[1], c.R_ref[1], mul(nw[0], c.R_ref[0])));
            n1 = fma_(nw[2], c.R_ref[5], fma_(nw[1], c.R_ref[4], mul(nw[0], c.R_ref[3])));
            n2 = fma_(nw[2], c.R_ref[8], fma_(nw[1], c.R_ref[7], mul(nw[0], c.R_ref[6])));
            cd = r.rz;
            allow = 1;
        }
    }
    a.cand_depth[g] = cd;
    a.cand_normal[3 * g] = n0; a.cand_normal[3 * g + 1] = n1; a.cand_normal[3 * g + 2] = n2;
    a.allow[g] = allow;
}

}  // namespace b200d
