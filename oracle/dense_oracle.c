/*
 * dense_oracle.c -- CPU restatement of the reference's Stage-4 dense hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The shipped path is the CUDA library under
 * 3d-reconstruction-tool_b200/csrc and has no CPU fallback.
 *
 * What it restates (reference = dackey-wav/3d-reconstruction-tool, paths relative
 * to /root/reference):
 *   src/core/dense_stereo.py:222-347   _plane_sweep_torch, _compute_ncc_torch
 *   src/core/dense_stereo.py:407-437   _backproject
 *   src/core/mvs_patchmatch.py:225-321 _patchmatch_cuda
 *   src/core/mvs_patchmatch.py:323-413 _compute_patch_cost, _ncc_cost
 *   src/core/mvs_patchmatch.py:415-534 _spatial_propagation, _random_refinement, _compute_confidence
 *   src/core/mvs_patchmatch.py:536-570 _fuse_depth_maps
 *
 * The arithmetic of that path lives in third-party wheels (PyTorch 2.11.0 CPU:
 * oneDNN 3.10.2 conv2d, ATen grid_sampler_2d, MKL sgemm), not in the reference
 * tree, and the reference has no tests or golden vectors.  The
 * rounding sequence below is the one SURVEY.md §7.4 measured for those library
 * calls; it is PINNED by tests/test_oracle_golden.py against fixtures produced
 * by executing the unmodified reference (oracle/make_golden.py): bit-exact NCC
 * volumes, cost maps, depth, confidence and masks.
 *
 * Rounding rules: every operation is one IEEE-754 binary32 round-to-nearest
 * operation; fused multiply-adds appear only where written as FMA().  Compile
 * with -ffp-contract=off and without -ffast-math (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FMA(a, b, c) __builtin_fmaf((a), (b), (c))
#define API __attribute__((visibility("default")))

/* Camera block shared by all entry points (all fp32, row-major 3x3):
 *   K     intrinsics at working resolution (K_scaled cast to fp32)
 *   Kinv  fp32 torch.inverse(K) computed by the caller (same bits on both sides)
 *   Rr,tr reference pose (world->camera);  Rs[S*9], ts[S*3] source poses          */
typedef struct {
    const float *K, *Kinv, *Rr, *tr, *Rs, *ts;
} cams_t;

API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

API void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- F.conv2d(x, ones(k,k)/(k*k), padding=k//2): row-major per-tap FMA chain, zero padding
 * (dense_stereo.py:325-341, mvs_patchmatch.py:397-408; SURVEY §7.4 row "F.conv2d"). */
API void orc_box_mean(const float *x, int H, int W, int k, float *out) {
    const int h = k / 2;
    const float w = 1.0f / (float)(k * k);
    const int Wp = W + 2 * h;
#pragma omp parallel
    {
        float *rows = (float *)calloc((size_t)Wp * k, sizeof(float));
        float *acc = (float *)malloc(sizeof(float) * W);
#pragma omp for schedule(static)
        for (int y = 0; y < H; ++y) {
            for (int i = 0; i < k; ++i) {
                int yy = y + i - h;
                float *dst = rows + (size_t)i * Wp;
                memset(dst, 0, sizeof(float) * Wp);
                if (yy >= 0 && yy < H) memcpy(dst + h, x + (size_t)yy * W, sizeof(float) * W);
            }
            for (int xx = 0; xx < W; ++xx) acc[xx] = 0.0f;
            for (int i = 0; i < k; ++i) {
                const float *r = rows + (size_t)i * Wp;
                for (int j = 0; j < k; ++j) {
                    const float *rj = r + j;
                    for (int xx = 0; xx < W; ++xx) acc[xx] = FMA(rj[xx], w, acc[xx]);
                }
            }
            memcpy(out + (size_t)y * W, acc, sizeof(float) * W);
        }
        free(rows);
        free(acc);
    }
}

/* ---- reference pixel (x,y) at depth d -> world point (mvs_patchmatch.py:341-347, dense_stereo.py:267-271)
 * torch.matmul with inner dim 3 = mul, fma, fma in index order. */
static inline void backproject_world(const cams_t *c, float x, float y, float d, float pw[3]) {
    float p[3];
    for (int n = 0; n < 3; ++n) {
        float a = x * c->Kinv[n * 3 + 0];
        a = FMA(y, c->Kinv[n * 3 + 1], a);
        a = FMA(1.0f, c->Kinv[n * 3 + 2], a);
        p[n] = a * d - c->tr[n];
    }
    for (int n = 0; n < 3; ++n) {
        float a = p[0] * c->Rr[0 * 3 + n];
        a = FMA(p[1], c->Rr[1 * 3 + n], a);
        a = FMA(p[2], c->Rr[2 * 3 + n], a);
        pw[n] = a;
    }
}

/* ---- world point -> source pixel (u,v) and depth z (mvs_patchmatch.py:351-359, dense_stereo.py:276-283) */
static inline void project_src(const cams_t *c, int s, const float pw[3], float *u, float *v, float *z) {
    const float *R = c->Rs + 9 * s, *t = c->ts + 3 * s;
    float q[3];
    for (int n = 0; n < 3; ++n) {
        float a = pw[0] * R[n * 3 + 0];
        a = FMA(pw[1], R[n * 3 + 1], a);
        a = FMA(pw[2], R[n * 3 + 2], a);
        q[n] = a + t[n];
    }
    *z = q[2];
    float zz = q[2] + 1e-8f;
    float px = q[0] / zz, py = q[1] / zz;
    float a = px * c->K[0];
    a = FMA(py, c->K[1], a);
    *u = a + c->K[2];
    a = px * c->K[3];
    a = FMA(py, c->K[4], a);
    *v = a + c->K[5];
}

/* ---- F.grid_sample(bilinear, zeros, align_corners=True) fed by the reference's normalisation
 * (mvs_patchmatch.py:367-377, dense_stereo.py:287-296): the pixel coordinate makes a round trip
 * through [-1,1]; taps outside the image read 0; out = mul, fma, fma, fma. */
static inline float sample_bilinear(const float *img, int H, int W, float u, float v) {
    float gx = (2.0f * u) / (float)(W - 1) - 1.0f;
    float gy = (2.0f * v) / (float)(H - 1) - 1.0f;
    float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1);
    float iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
    float x0 = floorf(ix), y0 = floorf(iy);
    float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    float wx1 = ix - x0, wx0 = x1 - ix, wy1 = iy - y0, wy0 = y1 - iy;
    float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
    int x0in = (x0 >= 0.0f) && (x0 <= (float)(W - 1));
    int x1in = (x1 >= 0.0f) && (x1 <= (float)(W - 1));
    int y0in = (y0 >= 0.0f) && (y0 <= (float)(H - 1));
    int y1in = (y1 >= 0.0f) && (y1 <= (float)(H - 1));
    float vnw = (x0in && y0in) ? img[(size_t)(int)y0 * W + (int)x0] : 0.0f;
    float vne = (x1in && y0in) ? img[(size_t)(int)y0 * W + (int)x1] : 0.0f;
    float vsw = (x0in && y1in) ? img[(size_t)(int)y1 * W + (int)x0] : 0.0f;
    float vse = (x1in && y1in) ? img[(size_t)(int)y1 * W + (int)x1] : 0.0f;
    float o = vnw * nw;
    o = FMA(vne, ne, o);
    o = FMA(vsw, sw, o);
    o = FMA(vse, se, o);
    return o;
}

/* ---- debug/test seam: per-pixel reprojection + sampling of one source at a per-pixel depth map
 * (the tensors the reference holds as proj[...,0], proj[...,1], z_src and src_sampled). */
API int orc_warp_sample(const float *src_img, int H, int W,
                        const float *K, const float *Kinv, const float *Rr, const float *tr,
                        const float *Rs, const float *ts, const float *depth,
                        float *samp, float *u_out, float *v_out, float *z_out) {
    cams_t c = {K, Kinv, Rr, tr, Rs, ts};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            float pw[3], u, v, z;
            backproject_world(&c, (float)x, (float)y, depth[i], pw);
            project_src(&c, 0, pw, &u, &v, &z);
            samp[i] = sample_bilinear(src_img, H, W, u, v);
            if (u_out) u_out[i] = u;
            if (v_out) v_out[i] = v;
            if (z_out) z_out[i] = z;
        }
    return 0;
}

/* Reference-side window statistics are source independent: mean1 and var1 = box(r*r) - mean1*mean1 */
typedef struct {
    int H, W, k;
    const float *ref;
    float *mean1, *var1;
    float *samp, *prod, *mean2, *m22, *m12; /* scratch (H*W each) */
} ncc_ws_t;

static void ws_init(ncc_ws_t *ws, const float *ref, int H, int W, int k) {
    size_t n = (size_t)H * W;
    ws->H = H; ws->W = W; ws->k = k; ws->ref = ref;
    ws->mean1 = (float *)malloc(n * sizeof(float));
    ws->var1 = (float *)malloc(n * sizeof(float));
    ws->samp = (float *)malloc(n * sizeof(float));
    ws->prod = (float *)malloc(n * sizeof(float));
    ws->mean2 = (float *)malloc(n * sizeof(float));
    ws->m22 = (float *)malloc(n * sizeof(float));
    ws->m12 = (float *)malloc(n * sizeof(float));
    orc_box_mean(ref, H, W, k, ws->mean1);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) ws->prod[i] = ref[i] * ref[i];
    orc_box_mean(ws->prod, H, W, k, ws->var1);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) ws->var1[i] = ws->var1[i] - ws->mean1[i] * ws->mean1[i];
}

static void ws_free(ncc_ws_t *ws) {
    free(ws->mean1); free(ws->var1); free(ws->samp); free(ws->prod);
    free(ws->mean2); free(ws->m22); free(ws->m12);
}

/* box statistics of the sampled image ws->samp against the reference: fills mean2, m22=box(s*s), m12=box(r*s) */
static void ws_src_stats(ncc_ws_t *ws) {
    size_t n = (size_t)ws->H * ws->W;
    orc_box_mean(ws->samp, ws->H, ws->W, ws->k, ws->mean2);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) ws->prod[i] = ws->samp[i] * ws->samp[i];
    orc_box_mean(ws->prod, ws->H, ws->W, ws->k, ws->m22);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) ws->prod[i] = ws->ref[i] * ws->samp[i];
    orc_box_mean(ws->prod, ws->H, ws->W, ws->k, ws->m12);
}

/* ================================================================== plane sweep
 * dense_stereo.py:262-310.  depths[D] are the fp32 casts of 1/linspace(1/dmax,1/dmin,D).
 * Outputs: depth_map, confidence (max count), best_idx (first index of the max);
 * ncc_dump (D,N,H,W) optional. */
API int orc_plane_sweep(const float *ref, const float *nbr_imgs /* N*H*W */, int N, int H, int W,
                        const float *K, const float *Kinv, const float *Rr, const float *tr,
                        const float *Rn, const float *tn, const float *depths, int D, int k, float thresh,
                        float *depth_map, float *confidence, int32_t *best_idx, float *ncc_dump) {
    cams_t c = {K, Kinv, Rr, tr, Rn, tn};
    ncc_ws_t ws;
    ws_init(&ws, ref, H, W, k);
    size_t n = (size_t)H * W;
    float *count = (float *)malloc(n * sizeof(float));
    float *zbuf = (float *)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; ++i) { confidence[i] = -1.0f; best_idx[i] = 0; }
    for (int d = 0; d < D; ++d) {
        const float depth = depths[d];
        memset(count, 0, n * sizeof(float));
        for (int s = 0; s < N; ++s) {
            const float *img = nbr_imgs + (size_t)s * n;
#pragma omp parallel for schedule(static)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float pw[3], u, v, z;
                    backproject_world(&c, (float)x, (float)y, depth, pw);
                    project_src(&c, s, pw, &u, &v, &z);
                    ws.samp[(size_t)y * W + x] = sample_bilinear(img, H, W, u, v);
                    zbuf[(size_t)y * W + x] = z;
                }
            ws_src_stats(&ws);
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) {
                float var2 = ws.m22[i] - ws.mean2[i] * ws.mean2[i];
                float cov = ws.m12[i] - ws.mean1[i] * ws.mean2[i];
                float denom = sqrtf(ws.var1[i] * var2 + 1e-8f);        /* dense_stereo.py:344 */
                float ncc = cov / denom;
                if (ncc_dump) ncc_dump[((size_t)d * N + s) * n + i] = ncc;
                if ((ncc > thresh) && (zbuf[i] > 0.1f)) count[i] += 1.0f;  /* :303-304 */
            }
        }
        for (size_t i = 0; i < n; ++i)                                   /* torch.max(dim=0): first max, :307 */
            if (count[i] > confidence[i]) { confidence[i] = count[i]; best_idx[i] = d; }
    }
    for (size_t i = 0; i < n; ++i) depth_map[i] = depths[best_idx[i]];
    free(count); free(zbuf);
    ws_free(&ws);
    return 0;
}

/* ================================================================== PatchMatch cost
 * mvs_patchmatch.py:323-390 with _ncc_cost :392-413.  mode 0: matching cost (strict bounds
 * half<=u<W-half, avg over valid sources, +inf if fewer than 2).  mode 1: confidence count
 * (:493-534: loose bounds 0<=u<W, ncc = 1 - cost > 0.6). */
static void pm_eval(ncc_ws_t *ws, const float *src_imgs, int S, const cams_t *c, const float *depth,
                    int mode, float *out, uint8_t *validbuf) {
    const int H = ws->H, W = ws->W, half = ws->k / 2;
    size_t n = (size_t)H * W;
    float *cnt = (float *)calloc(n, sizeof(float));
    for (size_t i = 0; i < n; ++i) out[i] = 0.0f;
    const float lo = mode ? 0.0f : (float)half;
    const float hix = mode ? (float)W : (float)(W - half);
    const float hiy = mode ? (float)H : (float)(H - half);
    for (int s = 0; s < S; ++s) {
        const float *img = src_imgs + (size_t)s * n;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t i = (size_t)y * W + x;
                float pw[3], u, v, z;
                backproject_world(c, (float)x, (float)y, depth[i], pw);
                project_src(c, s, pw, &u, &v, &z);
                validbuf[i] = (z > 0.1f) && (u >= lo) && (u < hix) && (v >= lo) && (v < hiy);
                ws->samp[i] = sample_bilinear(img, H, W, u, v);
            }
        ws_src_stats(ws);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; ++i) {
            float var2 = ws->m22[i] - ws->mean2[i] * ws->mean2[i];
            float cov = ws->m12[i] - ws->mean1[i] * ws->mean2[i];
            float ncc = cov / (sqrtf(ws->var1[i] * var2) + 1e-8f);     /* :410 */
            float cost = 1.0f - ncc;
            if (mode == 0) {
                out[i] = out[i] + (validbuf[i] ? cost : 0.0f);            /* :383 */
                cnt[i] = cnt[i] + (validbuf[i] ? 1.0f : 0.0f);
            } else {
                float ncc2 = 1.0f - cost;                                  /* :530 */
                if (validbuf[i] && (ncc2 > 0.6f)) out[i] += 1.0f;
            }
        }
    }
    if (mode == 0)
        for (size_t i = 0; i < n; ++i) {
            float avg = out[i] / (cnt[i] + 1e-8f);                         /* :387 */
            out[i] = (cnt[i] >= 2.0f) ? avg : INFINITY;
        }
    free(cnt);
}

API int orc_pm_cost(const float *ref, const float *src_imgs, int S, int H, int W,
                    const float *K, const float *Kinv, const float *Rr, const float *tr,
                    const float *Rs, const float *ts, const float *depth, int k, int mode, float *out) {
    cams_t c = {K, Kinv, Rr, tr, Rs, ts};
    ncc_ws_t ws;
    ws_init(&ws, ref, H, W, k);
    uint8_t *vb = (uint8_t *)malloc((size_t)H * W);
    pm_eval(&ws, src_imgs, S, &c, depth, mode, out, vb);
    free(vb);
    ws_free(&ws);
    return 0;
}

static void select_better(size_t n, const float *cost, const float *cd, const float *cn,
                          float *depth, float *normal, float *best) {
    for (size_t i = 0; i < n; ++i)
        if (cost[i] < best[i]) {                                           /* strict, NaN never wins (:452, :486) */
            depth[i] = cd[i];
            normal[3 * i] = cn[3 * i]; normal[3 * i + 1] = cn[3 * i + 1]; normal[3 * i + 2] = cn[3 * i + 2];
            best[i] = cost[i];
        }
}

/* ================================================================== full PatchMatch view
 * mvs_patchmatch.py:266-321.  depth/normal hold the injected initial hypotheses on entry and the
 * result on exit.  rand_u (iters*samples,H,W) and rand_n (iters*samples,H,W,3) are the injected
 * torch.rand / torch.randn draws of _random_refinement in call order.  depth_range_it[it] and
 * normal_range_it[it] are the fp32 casts of the Python-double expressions at :466-467.
 * costs_dump (iters*(2+samples),H,W) optional. */
API int orc_pm_view(const float *ref, const float *src_imgs, int S, int H, int W,
                    const float *K, const float *Kinv, const float *Rr, const float *tr,
                    const float *Rs, const float *ts, int k, int iters, int samples,
                    float depth_min, float depth_max, const float *depth_range_it, const float *normal_range_it,
                    const float *rand_u, const float *rand_n,
                    float *depth, float *normal, float *confidence, float *best_out, float *costs_dump) {
    cams_t c = {K, Kinv, Rr, tr, Rs, ts};
    ncc_ws_t ws;
    ws_init(&ws, ref, H, W, k);
    size_t n = (size_t)H * W;
    uint8_t *vb = (uint8_t *)malloc(n);
    float *best = (float *)malloc(n * sizeof(float));
    float *cd = (float *)malloc(n * sizeof(float));
    float *cn = (float *)malloc(3 * n * sizeof(float));
    float *cost = (float *)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; ++i) best[i] = INFINITY;
    int e = 0;
    for (int it = 0; it < iters; ++it) {
        /* _spatial_propagation (:415-457).  Even iterations: neighbour = (y+1,x) then (y,x+1);
         * odd: (y-1,x) then (y,x-1) -- the slicing, not the comments.
         * Border fill: depth_min, zero normal. */
        const int sgn = (it % 2 == 0) ? 1 : -1;
        for (int o = 0; o < 2; ++o) {
            const int dy = (o == 0) ? sgn : 0, dx = (o == 0) ? 0 : sgn;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    size_t i = (size_t)y * W + x;
                    int yy = y + dy, xx = x + dx;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                        size_t j = (size_t)yy * W + xx;
                        cd[i] = depth[j];
                        cn[3 * i] = normal[3 * j]; cn[3 * i + 1] = normal[3 * j + 1]; cn[3 * i + 2] = normal[3 * j + 2];
                    } else {
                        cd[i] = depth_min;
                        cn[3 * i] = cn[3 * i + 1] = cn[3 * i + 2] = 0.0f;
                    }
                }
            pm_eval(&ws, src_imgs, S, &c, cd, 0, cost, vb);
            if (costs_dump) memcpy(costs_dump + (size_t)(e) * n, cost, n * sizeof(float));
            ++e;
            select_better(n, cost, cd, cn, depth, normal, best);
        }
        /* _random_refinement (:459-491) */
        const float dr = depth_range_it[it], nr = normal_range_it[it];
        for (int sm = 0; sm < samples; ++sm) {
            const float *U = rand_u + (size_t)(it * samples + sm) * n;
            const float *N3 = rand_n + (size_t)(it * samples + sm) * n * 3;
            for (size_t i = 0; i < n; ++i) {
                float delta = (U[i] * 2.0f - 1.0f) * dr;                   /* :471 */
                float dn = depth[i] + delta;
                dn = dn < depth_min ? depth_min : dn;                      /* torch.clamp :472 */
                dn = dn > depth_max ? depth_max : dn;
                cd[i] = dn;
                float a = normal[3 * i] + N3[3 * i] * nr;                  /* :475-476 */
                float b = normal[3 * i + 1] + N3[3 * i + 1] * nr;
                float cc = normal[3 * i + 2] + N3[3 * i + 2] * nr;
                float nn = sqrtf(a * a + b * b + cc * cc);
                nn = nn < 1e-12f ? 1e-12f : nn;                            /* F.normalize eps */
                cn[3 * i] = a / nn; cn[3 * i + 1] = b / nn; cn[3 * i + 2] = cc / nn;
            }
            pm_eval(&ws, src_imgs, S, &c, cd, 0, cost, vb);
            if (costs_dump) memcpy(costs_dump + (size_t)(e) * n, cost, n * sizeof(float));
            ++e;
            select_better(n, cost, cd, cn, depth, normal, best);
        }
    }
    pm_eval(&ws, src_imgs, S, &c, depth, 1, confidence, vb);               /* :311 */
    if (best_out) memcpy(best_out, best, n * sizeof(float));
    free(vb); free(best); free(cd); free(cn); free(cost);
    ws_free(&ws);
    return 0;
}

/* ================================================================== mask + back-projection
 * mode 0: PatchMatch fusion, valid = conf >= min_conf           (mvs_patchmatch.py:545-565)
 * mode 1: plane sweep,      valid = conf >= min_conf && depth>0 (dense_stereo.py:414-437)
 * Row-major np.where order.  Kinv64 = np.linalg.inv(K_scaled) (f64), R,t f64.
 * points (M,3) f64 = ((Kinv [x,y,1]) * d - t) @ R ; colors (M,3) u8 = BGR reversed.
 * The reference evaluates this through NumPy/BLAS in f64 whose summation order is unspecified,
 * so coordinates are compared with a tolerance; mask, count, order and colours are exact. */
API int64_t orc_mask_backproject(const float *depth, const float *conf, const uint8_t *bgr, int H, int W,
                                 const double *Kinv64, const double *R, const double *t, float min_conf, int mode,
                                 double *points, uint8_t *colors) {
    int64_t m = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            int ok = conf[i] >= min_conf;
            if (mode == 1) ok = ok && (depth[i] > 0.0f);
            if (!ok) continue;
            if (points) {
                double p[3];
                for (int n = 0; n < 3; ++n) {
                    double r = (double)x * Kinv64[n * 3] + (double)y * Kinv64[n * 3 + 1] + Kinv64[n * 3 + 2];
                    p[n] = r * (double)depth[i] - t[n];
                }
                for (int n = 0; n < 3; ++n)
                    points[3 * m + n] = p[0] * R[0 * 3 + n] + p[1] * R[1 * 3 + n] + p[2] * R[2 * 3 + n];
                colors[3 * m + 0] = bgr[3 * i + 2];
                colors[3 * m + 1] = bgr[3 * i + 1];
                colors[3 * m + 2] = bgr[3 * i + 0];
            }
            ++m;
        }
    return m;
}

/* ================================================================== cross-view EXTENSIONS
 * The reference has no geometric consistency check (its check is photometric) and no view
 * propagation (README.md:117 names it, the code has none).  BASELINE.json's north_star asks for
 * both, so the repo DEFINES them; there is no dackey-wav behaviour to be faithful to ("parity unpinned").
 * These two functions are that definition, written with the same fp32 primitives as the pinned path
 * (backproject / project sequences above), so the CUDA kernels (csrc/geo_kernels.cuh) can be checked
 * bit for bit against it. */

static inline void backproject_cam(const float *Kinv, const float *R, const float *t, float x, float y, float d,
                                   float pw[3]) {
    float p[3];
    for (int n = 0; n < 3; ++n) {
        float a = x * Kinv[n * 3 + 0];
        a = FMA(y, Kinv[n * 3 + 1], a);
        a = a + Kinv[n * 3 + 2];
        p[n] = a * d - t[n];
    }
    for (int n = 0; n < 3; ++n) {
        float a = p[0] * R[0 * 3 + n];
        a = FMA(p[1], R[1 * 3 + n], a);
        a = FMA(p[2], R[2 * 3 + n], a);
        pw[n] = a;
    }
}

static inline void world_to_cam(const float *R, const float *t, const float pw[3], float q[3]) {
    for (int n = 0; n < 3; ++n) {
        float a = pw[0] * R[n * 3 + 0];
        a = FMA(pw[1], R[n * 3 + 1], a);
        a = FMA(pw[2], R[n * 3 + 2], a);
        q[n] = a + t[n];
    }
}

static inline void to_pixel(const float *K, float px, float py, float *u, float *v) {
    float a = px * K[0];
    a = FMA(py, K[1], a);
    *u = a + K[2];
    a = px * K[3];
    a = FMA(py, K[4], a);
    *v = a + K[5];
}

typedef struct { int ok, us, vs; float rz, du, dv; } round_trip_t;

/* reference pixel (x,y) at depth d -> source s (nearest pixel, its depth) -> back into the reference camera */
static inline round_trip_t round_trip(const cams_t *c, int s, const float *depth_src, int W, int H, int x, int y,
                                      const float Xw[3]) {
    round_trip_t r = {0, 0, 0, 0.0f, 0.0f, 0.0f};
    const float *R = c->Rs + 9 * s, *t = c->ts + 3 * s;
    float q[3];
    world_to_cam(R, t, Xw, q);
    if (!(q[2] > 1e-6f)) return r;
    float zz = q[2] + 1e-8f, u, v;
    to_pixel(c->K, q[0] / zz, q[1] / zz, &u, &v);
    float uf = rintf(u), vf = rintf(v);                                   /* round half to even */
    if (!((uf >= 0.0f) && (uf <= (float)(W - 1)) && (vf >= 0.0f) && (vf <= (float)(H - 1)))) return r;
    r.us = (int)uf; r.vs = (int)vf;
    float ds = depth_src[(size_t)r.vs * W + r.us];
    if (!(ds > 0.0f)) return r;
    float Xb[3], pr[3];
    backproject_cam(c->Kinv, R, t, uf, vf, ds, Xb);
    world_to_cam(c->Rr, c->tr, Xb, pr);
    if (!(pr[2] > 1e-6f)) return r;
    float zr = pr[2] + 1e-8f, pu, pv;
    to_pixel(c->K, pr[0] / zr, pr[1] / zr, &pu, &pv);
    r.rz = pr[2];
    r.du = pu - (float)x;
    r.dv = pv - (float)y;
    r.ok = 1;
    return r;
}

/* count[p] = number of sources s whose depth map agrees with depth_ref at p:
 * |p' - p|^2 < px_thresh^2 and |r.z - d| < rel_thresh * d  (see csrc/geo_kernels.cuh header). */
API int orc_geo_consistency(const float *depth_ref, const float *depth_srcs /* S*H*W */, int S, int H, int W,
                            const float *K, const float *Kinv, const float *Rr, const float *tr,
                            const float *Rs, const float *ts, float px_thresh, float rel_thresh, float *count) {
    cams_t c = {K, Kinv, Rr, tr, Rs, ts};
    size_t n = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            float d = depth_ref[i], cnt = 0.0f;
            if (d > 0.0f) {
                float Xw[3];
                backproject_cam(c.Kinv, c.Rr, c.tr, (float)x, (float)y, d, Xw);
                float px2 = px_thresh * px_thresh, dtol = rel_thresh * d;
                for (int s = 0; s < S; ++s) {
                    round_trip_t r = round_trip(&c, s, depth_srcs + (size_t)s * n, W, H, x, y, Xw);
                    int ok = r.ok && (r.du * r.du + r.dv * r.dv < px2) && (fabsf(r.rz - d) < dtol);
                    cnt = cnt + (ok ? 1.0f : 0.0f);
                }
            }
            count[i] = cnt;
        }
    return 0;
}

/* view-propagation proposal of source s (index into the camera block) for every reference pixel:
 * candidate depth = r.z if the round trip succeeds and depth_min <= r.z <= depth_max, candidate normal =
 * R_ref R_s^T n_s[p_s]; otherwise the pixel keeps its hypothesis and allow = 0. */
API int orc_viewprop_candidates(const float *depth, const float *normal, const float *depth_src, const float *normal_src,
                                int s, int H, int W, const float *K, const float *Kinv, const float *Rr, const float *tr,
                                const float *Rs, const float *ts, float depth_min, float depth_max,
                                float *cand_depth, float *cand_normal, uint8_t *allow) {
    cams_t c = {K, Kinv, Rr, tr, Rs, ts};
    const float *R = Rs + 9 * s;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            float d = depth[i];
            float cd = d, n0 = normal[3 * i], n1 = normal[3 * i + 1], n2 = normal[3 * i + 2];
            uint8_t al = 0;
            if (d > 0.0f) {
                float Xw[3];
                backproject_cam(c.Kinv, c.Rr, c.tr, (float)x, (float)y, d, Xw);
                round_trip_t r = round_trip(&c, s, depth_src, W, H, x, y, Xw);
                if (r.ok && (r.rz >= depth_min) && (r.rz <= depth_max)) {
                    const float *ns = normal_src + 3 * ((size_t)r.vs * W + r.us);
                    float nw[3];
                    for (int n = 0; n < 3; ++n) nw[n] = FMA(ns[2], R[2 * 3 + n], FMA(ns[1], R[1 * 3 + n], ns[0] * R[0 * 3 + n]));
                    n0 = FMA(nw[2], Rr[2], FMA(nw[1], Rr[1], nw[0] * Rr[0]));
                    n1 = FMA(nw[2], Rr[5], FMA(nw[1], Rr[4], nw[0] * Rr[3]));
                    n2 = FMA(nw[2], Rr[8], FMA(nw[1], Rr[7], nw[0] * Rr[6]));
                    cd = r.rz;
                    al = 1;
                }
            }
            cand_depth[i] = cd;
            cand_normal[3 * i] = n0; cand_normal[3 * i + 1] = n1; cand_normal[3 * i + 2] = n2;
            allow[i] = al;
        }
    return 0;
}
