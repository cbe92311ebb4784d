// geo_kernels.cuh -- cross-view geometric consistency count (EXTENSION: no dackey-wav oracle).
//
// The reference's "multi-view consistency" is photometric (`_compute_confidence` counts sources with
// NCC > 0.6).  BASELINE.json's north_star additionally asks for a geometric filter fed by the depth maps of the
// other views (all-gathered over NCCL when views are sharded).  This kernel is that filter, in its usual
// forward-backward form: for reference pixel p with depth d,
//     X   = backproject_ref(p, d)                       world point
//     p_s = project_src(X)  ->  nearest source pixel, its depth d_s from the source's depth map
//     X'  = backproject_src(p_s, d_s) ;  (p', d') = project_ref(X')
//     consistent_s  <=>  |p - p'| < px_thresh  and  |d' - d| / d < rel_thresh
// out[p] = number of consistent sources.  fp32, one thread per pixel, coalesced, HBM/L2-bound.
// Checked against a PyTorch restatement of the same definition (tests/test_gpu_geo.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200dense.h"

namespace b200d {

struct GeoArgs {
    const float *depth_ref;
    const float *depth_src[B200D_MAX_SRC];
    b200d_cams cams;
    int H, W;
    float px_thresh, rel_thresh;
    float *out;
};

__global__ void __launch_bounds__(256) geo_consistency_kernel(const __grid_constant__ GeoArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const b200d_cams &c = a.cams;
    const size_t g = (size_t)y * a.W + x;
    const float d = a.depth_ref[g];
    const float fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
    // camera-frame point of the reference pixel, then world: Xw = R_ref^T (Xc - t_ref)
    const float pc[3] = {(x - cx) / fx * d - c.t_ref[0], (y - cy) / fy * d - c.t_ref[1], d - c.t_ref[2]};
    float Xw[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) Xw[n] = pc[0] * c.R_ref[0 * 3 + n] + pc[1] * c.R_ref[1 * 3 + n] + pc[2] * c.R_ref[2 * 3 + n];
    float count = 0.0f;
    for (int s = 0; s < c.n_src; ++s) {
        const float *R = c.R_src[s], *t = c.t_src[s];
        float q[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) q[n] = Xw[0] * R[n * 3] + Xw[1] * R[n * 3 + 1] + Xw[2] * R[n * 3 + 2] + t[n];
        if (!(q[2] > 1e-6f) || !(d > 0.0f)) continue;
        const int us = __float2int_rn(fx * q[0] / q[2] + cx), vs = __float2int_rn(fy * q[1] / q[2] + cy);
        if (us < 0 || us >= a.W || vs < 0 || vs >= a.H) continue;
        const float ds = __ldg(a.depth_src[s] + (size_t)vs * a.W + us);
        if (!(ds > 0.0f)) continue;
        // back-project the source pixel at its own depth, return to the reference camera
        const float qs[3] = {(us - cx) / fx * ds - t[0], (vs - cy) / fy * ds - t[1], ds - t[2]};
        float Xb[3], pr[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) Xb[n] = qs[0] * R[0 * 3 + n] + qs[1] * R[1 * 3 + n] + qs[2] * R[2 * 3 + n];
#pragma unroll
        for (int n = 0; n < 3; ++n)
            pr[n] = Xb[0] * c.R_ref[n * 3] + Xb[1] * c.R_ref[n * 3 + 1] + Xb[2] * c.R_ref[n * 3 + 2] + c.t_ref[n];
        if (!(pr[2] > 1e-6f)) continue;
        const float du = fx * pr[0] / pr[2] + cx - x, dv = fy * pr[1] / pr[2] + cy - y;
        const bool ok = (du * du + dv * dv < a.px_thresh * a.px_thresh) && (fabsf(pr[2] - d) < a.rel_thresh * d);
        count += ok ? 1.0f : 0.0f;
    }
    a.out[g] = count;
}

}  // namespace b200d
