// sample_kernels.cuh -- hypothesis generation + reprojection + bilinear sampling, one thread per pixel.
//
// First half of a cost evaluation (reference: src/core/mvs_patchmatch.py:341-377 and
// src/core/dense_stereo.py:267-296).  For every reference pixel the kernel
//   1. forms the candidate depth of this evaluation
//        C_MAP     the given depth map                        (_compute_patch_cost / _compute_confidence)
//        C_SHIFT   the neighbour's depth, depth_min at the border   (_spatial_propagation :429-442)
//        C_REFINE  clamp(depth + (2U-1) * range)                    (_random_refinement :471-472)
//        C_PLANE   one fronto-parallel plane                        (_plane_sweep_torch :265-271)
//   2. back-projects it once (shared by all sources), and for each source view reprojects,
//      tests validity and samples the source image (grid_sample semantics),
//   3. writes the S sampled images (pitch-padded, so the NCC kernel can fetch tiles with TMA),
//      the candidate depth and a per-pixel validity bit mask.
// There is no halo and no redundancy here: every (pixel, source) sample is computed exactly once;
// the windowed statistics over these sampled images are the NCC kernel's job (ncc_kernels.cuh).
// The sampled images live in an L2-sized scratch (S x H x W fp32) and are consumed immediately.
#pragma once
#include "device_math.cuh"

namespace b200d {

enum CandMode { C_MAP = 0, C_SHIFT = 1, C_REFINE = 2, C_PLANE = 3 };
enum BoundsMode { B_PATCH = 0, B_IMAGE = 1, B_DEPTH = 2 };   // validity test of a projection

struct SampleArgs {
    const float *src[B200D_MAX_SRC];
    b200d_cams cams;
    SrcPairCams pc[B200D_MAX_SRC / 2];   // R_src / t_src of sources (2i, 2i+1) interleaved (fill_pair_cams)
    int H, W, Wp, pad;                   // sampled maps: row pitch Wp floats, pixel x at column x + pad
    size_t map_stride;                   // floats between consecutive sampled maps
    const float *depth_in;               // (H,W) current depth (C_MAP / C_SHIFT / C_REFINE)
    int dy, dx;                          // C_SHIFT offset
    float depth_min, depth_max, dr;      // fill / clamp / refinement range
    float plane_depth;                   // C_PLANE
    const float *rand_u;                 // injected U (H,W) of this evaluation, or NULL -> Philox
    uint64_t seed;
    uint32_t view_id, eval_id;
    int half;                            // patch/2 (B_PATCH bounds)
    float *smap;                         // out: S sampled maps
    float *cand;                         // out: (H,W) candidate depth (may be NULL for C_PLANE)
    uint8_t *valid;                      // out: (H,W) bit s = projection into source s is valid
};

constexpr int SBX = 32, SBY = 8;         // thread block: 32 x 8 pixels

// Sources can be processed two at a time through the packed fp32x2 path (project_sample_fast2; an odd count leaves one
// source to the scalar path; results are bit-identical either way).  Measured on B200 (profiles/microbench/
// r02_pairs_cmd.sh, r02_pairs_ab2_cmd.sh): the plane sweep's sampling kernel is issue-bound and gains 12 % (56 -> 49 us
// per plane, sweep 19.9 -> 21.1 Gpix*planes/s); the PatchMatch sampling kernels gain 2-4 % in isolation, but the 4-stream
// PatchMatch step is bound by FP32-pipe cycles across kernels, where an FFMA2 counts twice, and loses 0.7-1.5 %
// (181.0 -> 178.4 / 179.6 views/s, reproducible).  So: B200D_SAMPLE_PAIRS = 1 (default) packs the plane sweep only,
// 2 packs every mode, 0 none (A/B builds).
#ifndef B200D_SAMPLE_PAIRS
#define B200D_SAMPLE_PAIRS 1
#endif
#ifndef B200D_SAMPLE_MIN_BLOCKS
#define B200D_SAMPLE_MIN_BLOCKS 6
#endif

inline void fill_pair_cams(SrcPairCams *pc, const b200d_cams &c) {
    for (int i = 0; i < B200D_MAX_SRC / 2; ++i) {
        for (int j = 0; j < 9; ++j) pc[i].R[j] = make_float2(c.R_src[2 * i][j], c.R_src[2 * i + 1][j]);
        for (int j = 0; j < 3; ++j) pc[i].t[j] = make_float2(c.t_src[2 * i][j], c.t_src[2 * i + 1][j]);
    }
}

// NS > 0: number of sources known at compile time -- the loop is fully unrolled and every camera constant becomes a
// constant-bank operand of the arithmetic instruction that uses it (no uniform loads, no index arithmetic);
// NS == 0: run-time count.
template <int CAND, int BOUNDS, int NS>
__global__ void __launch_bounds__(SBX *SBY, B200D_SAMPLE_MIN_BLOCKS) sample_kernel(const __grid_constant__ SampleArgs a) {
    const int x = blockIdx.x * SBX + threadIdx.x, y = blockIdx.y * SBY + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    const int W = a.W, H = a.H;
    const size_t g = (size_t)y * W + x;

    float cd;
    if (CAND == C_MAP) {
        cd = __ldg(a.depth_in + g);
    } else if (CAND == C_SHIFT) {
        const int nx = x + a.dx, ny = y + a.dy;
        cd = (nx >= 0 && nx < W && ny >= 0 && ny < H) ? __ldg(a.depth_in + (size_t)ny * W + nx) : a.depth_min;
    } else if (CAND == C_REFINE) {
        float u;
        if (a.rand_u) {
            u = __ldg(a.rand_u + g);
        } else {
            const uint4 rn = philox4x32_10(make_uint4((uint32_t)g, a.eval_id, a.view_id, 0u),
                                           make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
            u = u01(rn.x);
        }
        const float delta = mul(sub(mul(u, 2.0f), 1.0f), a.dr);
        float dn = add(__ldg(a.depth_in + g), delta);
        dn = dn < a.depth_min ? a.depth_min : dn;
        dn = dn > a.depth_max ? a.depth_max : dn;
        cd = dn;
    } else {
        cd = a.plane_depth;
    }

    float pw[3];
    backproject_world(a.cams, (float)x, (float)y, cd, pw);
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float rw = rcp_refined(wm1), rh = rcp_refined(hm1);
    const float lo = (BOUNDS == B_PATCH) ? (float)a.half : 0.0f;
    const float hix = (BOUNDS == B_PATCH) ? (float)(W - a.half) : (float)W;
    const float hiy = (BOUNDS == B_PATCH) ? (float)(H - a.half) : (float)H;

    unsigned vmask = 0;
    float *out = a.smap + (size_t)y * a.Wp + x + a.pad;
    auto one_source = [&](int s) {
        float u, v, z, smp;
        if (!project_sample_fast(a.cams, s, pw, a.src[s], W, H, wm1, hm1, rw, rh, u, v, z, smp)) {
            float uv[2];
            smp = project_sample_exact(a.cams, s, pw, a.src[s], W, H, wm1, hm1, uv);
            u = uv[0];
            v = uv[1];
        }
        bool ok = z > 0.1f;
        if (BOUNDS != B_DEPTH) ok = ok & (u >= lo) & (u < hix) & (v >= lo) & (v < hiy);
        vmask |= (unsigned)ok << s;
        out[(size_t)s * a.map_stride] = smp;
    };
    auto two_sources = [&](int s) {          // sources s and s + 1 (s even)
        float u[2], v[2], z[2], smp[2];
        bool trusted[2];
        project_sample_fast2(a.cams, a.pc[s >> 1], pw, a.src[s], a.src[s + 1], W, H, wm1, hm1, rw, rh, u, v, z, smp, trusted);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!trusted[h]) {
                float uv[2];
                smp[h] = project_sample_exact(a.cams, s + h, pw, a.src[s + h], W, H, wm1, hm1, uv);
                u[h] = uv[0];
                v[h] = uv[1];
            }
            bool ok = z[h] > 0.1f;
            if (BOUNDS != B_DEPTH) ok = ok & (u[h] >= lo) & (u[h] < hix) & (v[h] >= lo) & (v[h] < hiy);
            vmask |= (unsigned)ok << (s + h);
            out[(size_t)(s + h) * a.map_stride] = smp[h];
        }
    };
    constexpr bool PAIRS = B200D_SAMPLE_PAIRS == 2 || (B200D_SAMPLE_PAIRS == 1 && CAND == C_PLANE);
    if (NS > 0) {
#pragma unroll
        for (int s = 0; s + 1 < NS; s += 2) {
            if (PAIRS) two_sources(s);
            else { one_source(s); one_source(s + 1); }
        }
        if (NS & 1) one_source(NS - 1);
    } else {
        const int S = a.cams.n_src;
        int s = 0;
        if (PAIRS)
            for (; s + 1 < S; s += 2) two_sources(s);
#pragma unroll 2
        for (; s < S; ++s) one_source(s);
    }
    if (a.cand) a.cand[g] = cd;
    a.valid[g] = (uint8_t)vmask;
}

}  // namespace b200d
