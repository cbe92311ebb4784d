/*
 * b200dense.h -- C ABI of the B200-native dense-reconstruction hot path.
 *
 * Drop-in boundary for Stage 4 of dackey-wav/3d-reconstruction-tool.  The reference has no
 * FFI of its own: its boundary is the Python class API, and every device
 * operation is a PyTorch eager call.  Each entry point below replaces one reference method
 * (cited as /root/reference-relative file:line) and is what a ctypes binding inside those
 * methods would call (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain C, no torch / C++ types; all image-like buffers are row-major, contiguous.
 *   - pointers named d_* are DEVICE pointers (sm_100a), everything else is host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *     enqueue work; they never synchronise unless stated.  Thread-safe per stream.
 *   - return value: 0 on success, negative b200d_status on error; b200d_last_error() gives the
 *     message for the calling thread.  No exceptions cross the boundary, nothing is printed.
 *   - caller owns every buffer; the library allocates nothing unless stated.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     B200D_ERR_CUDA.
 *
 * Arithmetic modes
 *   B200D_PARITY  IEEE-RN fp32 in the exact operation order of the reference's CPU path
 *                 (SURVEY.md §7.4): bit-identical to oracle/dense_oracle.c.
 *                 In the plane sweep, which uses the NCC only through `ncc > thresh`, the comparison is decided
 *                 from separable window sums plus a rigorous error bound and only undecided outputs redo the
 *                 reference's chains: same counts / arg-max bit for bit (csrc/fast_kernels.cuh).
 *   B200D_PARITY_CHAINS  same results as B200D_PARITY; the plane sweep evaluates the reference's chains for
 *                 every output (the A/B twin of the bounded decision, used by tests and profiling).
 *   B200D_PARITY_CHECK   diagnostic twin of B200D_PARITY for the plane sweep: same results, and for every output the
 *                 observed distance between the separable sums and the reference's chains is compared with the bound
 *                 the decision relies on (b200d_sweep_bound_check).  Slow; tests only.
 *   B200D_FAST    OPT-IN, NOT BIT-EXACT, never the default: composed reprojection, direct bilinear lookup,
 *                 separable window sums, MUFU reciprocal / rsqrt.  It misses the north_star tolerances wherever a
 *                 patch is flat (var = E[x^2] - E[x]^2 is ill-conditioned) and its agreement with
 *                 the oracle is measured and reported, never asserted as parity (tests/test_gpu_fast.py).
 */
#ifndef B200DENSE_H
#define B200DENSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200D_API __attribute__((visibility("default")))
#else
#define B200D_API
#endif

#define B200D_MAX_SRC 8          /* source / neighbour views per reference view */
#define B200D_ABI_VERSION 3

typedef enum {
    B200D_OK = 0,
    B200D_ERR_ARG = -1,          /* bad pointer / size / unsupported patch size */
    B200D_ERR_CUDA = -2,         /* CUDA runtime error (message in b200d_last_error) */
    B200D_ERR_UNSUPPORTED = -3
} b200d_status;

typedef enum { B200D_PARITY = 0, B200D_FAST = 1, B200D_PARITY_CHAINS = 2, B200D_PARITY_CHECK = 3 } b200d_arith;

/* Camera block, fp32, exactly what the reference uploads per reference view
 * (mvs_patchmatch.py:237-257, dense_stereo.py:232-259): K = K_scaled cast to fp32,
 * Kinv = fp32 inverse of K (torch.inverse on the host), R/t world->camera. */
typedef struct {
    float K[9];
    float Kinv[9];
    float R_ref[9];
    float t_ref[3];
    int32_t n_src;
    float R_src[B200D_MAX_SRC][9];
    float t_src[B200D_MAX_SRC][3];
} b200d_cams;

/* Random numbers for PatchMatch.  Either injected tensors (parity tests: the reference's own
 * torch.rand / torch.randn draws, call order of mvs_patchmatch.py:271,279-280,471,475) or a
 * counter-based Philox4x32-10 stream keyed by (seed, view, evaluation, pixel). */
typedef struct {
    uint64_t seed;               /* used when the matching pointer is NULL */
    uint32_t view_id;
    const float *d_init_depth;   /* (H,W)   injected initial depth, or NULL */
    const float *d_init_normal;  /* (H,W,3) injected initial normal, or NULL */
    const float *d_rand_u;       /* (iters*samples,H,W)   uniform [0,1) draws, or NULL */
    const float *d_rand_n;       /* (iters*samples,H,W,3) standard normal draws, or NULL */
} b200d_rng;

/* PatchMatch parameters (PatchMatchMVS.__init__, mvs_patchmatch.py:43-59). */
typedef struct {
    int32_t patch;               /* odd NCC window: 3,5,7,9,11,13 */
    int32_t iters;
    int32_t samples;
    int32_t arith;               /* b200d_arith */
    /* f64 like the Python attributes: the per-iteration ranges (dmax-dmin)*0.5^it and 0.5*0.5^it
     * are evaluated in double and then cast to fp32 (mvs_patchmatch.py:466-467); clamp and border
     * fill use the fp32 casts. */
    double depth_min, depth_max;
    /* Propagation schedule.  B200D_SCHEDULE_REFERENCE: the reference's two whole-image shifted-neighbour evaluations
     * per iteration (mvs_patchmatch.py:415-457) -- the graded, bit-exact mode.  B200D_SCHEDULE_REDBLACK: EXTENSION with
     * no counterpart in the reference: per iteration, for each checkerboard colour in turn, the hypotheses of
     * the four axis neighbours are evaluated and only pixels of that colour may adopt them (8 evaluations), so the
     * second colour already sees the first colour's updates; refinement is unchanged. */
    int32_t schedule;
    /* Staged execution, for schedules that interleave the views between iterations (view propagation, an EXTENSION):
     * run iterations [it_begin, it_end) only (it_end == 0 means `iters`), optionally without the initialisation
     * (B200D_STAGE_NO_INIT: d_depth / d_normal / d_best already hold the state) and without the final confidence
     * pass (B200D_STAGE_NO_CONF).  All zero = the whole view, which is the reference's `_patchmatch_cuda`. */
    int32_t stage;
    int32_t it_begin, it_end;
} b200d_pm_params;

typedef enum { B200D_SCHEDULE_REFERENCE = 0, B200D_SCHEDULE_REDBLACK = 1 } b200d_schedule;
typedef enum { B200D_STAGE_NO_INIT = 1, B200D_STAGE_NO_CONF = 2 } b200d_stage;

/* Device scratch of one view in flight, shared by all compute entry points below:
 * reference window statistics, candidate depth, validity bits, a pitch-padded copy of the reference
 * (TMA source when W is not a multiple of 4) and B200D_MAX_SRC pitch-padded sampled images
 * (~65 bytes per pixel).  Calls that may run concurrently (different streams) need different scratch. */
B200D_API size_t b200d_scratch_bytes(int H, int W);

B200D_API int b200d_abi_version(void);
B200D_API const char *b200d_last_error(void);
B200D_API int b200d_device_count(void);

/* Reference-side window statistics mean1 = box(r), var1 = box(r*r) - mean1^2
 * (first two conv2d of _ncc_cost / _compute_ncc_torch, mvs_patchmatch.py:403-406,
 * dense_stereo.py:333-337; constant per reference view). */
B200D_API int b200d_ref_stats(const float *d_ref, int H, int W, int patch, float *d_mean1, float *d_var1,
                              void *d_scratch, void *stream);

/* Replaces PatchMatchMVS._compute_patch_cost (mvs_patchmatch.py:323-390) when mode==0 and
 * PatchMatchMVS._compute_confidence (:493-534) when mode==1, for a given per-pixel depth map.
 * d_src[s] are the source grays (H,W) fp32; d_out (H,W) receives the cost (may hold +inf / NaN) or the
 * count of consistent sources. */
B200D_API int b200d_pm_cost(const float *d_ref, const float *const *d_src, const b200d_cams *cams,
                            const float *d_depth, int H, int W, int patch, int mode, int arith, float *d_out,
                            void *d_scratch, void *stream);

/* Replaces PatchMatchMVS._patchmatch_cuda (mvs_patchmatch.py:225-321) from initialisation to the
 * confidence map: init -> iters x (2 shifted-neighbour evaluations + `samples` random
 * refinements, strict `<` selection) -> confidence.  Results: d_depth (H,W), d_normal (H,W,3),
 * d_conf (H,W), optional d_best (H,W); d_scratch: b200d_scratch_bytes(H,W) bytes.  d_costs (iters*(P+samples),H,W) optional dump of every
 * evaluated cost map (tests), P = 2 propagation evaluations per iteration in the reference schedule, 8 in red-black. */
B200D_API int b200d_pm_view(const float *d_ref, const float *const *d_src, const b200d_cams *cams,
                  int H, int W, const b200d_pm_params *prm, const b200d_rng *rng,
                  void *d_scratch, float *d_depth, float *d_normal, float *d_conf, float *d_best,
                  float *d_costs, void *stream);

/* EXTENSION (the reference's README lists "view propagation", README.md:117, but its code has none -- SURVEY F2; no
 * dackey-wav oracle): one propagation step from the source views' current hypotheses.  For every source s in turn:
 * each reference pixel p at its current depth d is projected into s, the nearest pixel's depth d_s of s's depth map is
 * back-projected and brought into the reference camera; its depth there (if inside [depth_min, depth_max]) with the
 * normal R_ref R_s^T n_s is p's candidate, costed like any other hypothesis map and adopted where `cost < best`
 * (pixels whose round trip fails keep their hypothesis and may not adopt).  d_src_depth[s] (H,W) and d_src_normal[s]
 * (H,W,3) are SNAPSHOTS of the sources' maps (all-gathered over NCCL when the views are sharded); d_depth, d_normal,
 * d_best are updated in place; d_costs (S,H,W) optional dump.  Definition restated in oracle/dense_oracle.c
 * (orc_viewprop_candidates) + tests/test_gpu_viewprop.py; kernels csrc/geo_kernels.cuh. */
B200D_API int b200d_pm_viewprop(const float *d_ref, const float *const *d_src, const b200d_cams *cams, int H, int W,
                                const b200d_pm_params *prm, const float *const *d_src_depth,
                                const float *const *d_src_normal, void *d_scratch, float *d_depth, float *d_normal,
                                float *d_best, float *d_costs, void *stream);

/* Replaces DenseStereoReconstructor._plane_sweep_torch (dense_stereo.py:222-316):
 * for each plane d (given order) and neighbour n: reproject at constant depth, bilinear sample,
 * box NCC (_compute_ncc_torch :318-347), count (ncc > thresh) & (z > 0.1); first arg-max over d.
 * depths (D) fp32 in HOST memory (each plane is a kernel argument).  Outputs d_depth_map, d_conf (H,W)
 * fp32, d_idx (H,W) int32; d_ncc (D,N,H,W) optional dump of the NCC volume (tests).
 * d_scratch: b200d_scratch_bytes(H,W) bytes. */
B200D_API int b200d_plane_sweep(const float *d_ref, const float *const *d_nbr, const b200d_cams *cams,
                                const float *depths, int D, int H, int W, int patch, float thresh, int arith,
                                float *d_depth_map, float *d_conf, int32_t *d_idx, float *d_ncc, void *d_scratch,
                                void *stream);

/* Diagnostic (synchronises the device): number of (pixel, neighbour, plane) outputs of the B200D_PARITY plane sweep
 * whose bounded decision was inconclusive and which were redone with the reference's chains, since the last reset. */
B200D_API int b200d_sweep_fallbacks(uint64_t *count, int reset);

/* Diagnostic (synchronises the device): after plane sweeps run with B200D_PARITY_CHECK, ratios3 (HOST, 3 floats) receives
 * the largest observed |separable sum - reference chain| / bound for the three window sums box(s), box(s*s), box(r*s) over
 * every output evaluated since the last reset.  Values <= 1 confirm the error bound behind the bit-exact decision. */
B200D_API int b200d_sweep_bound_check(float *ratios3, int reset);

/* Replaces the per-view body of PatchMatchMVS._fuse_depth_maps (mvs_patchmatch.py:545-565, mode 0:
 * valid = conf >= min_conf) and DenseStereoReconstructor._backproject (dense_stereo.py:414-437,
 * mode 1: valid = conf >= min_conf && depth > 0): mask, row-major ordered compaction,
 * back-projection X = ((Kinv [x,y,1]) d - t) R in f64, colour BGR->RGB.
 * Kinv64/R64/t64 are HOST f64.  d_points (>= H*W*3 f64), d_colors (>= H*W*3 u8),
 * d_count (1 x int64), d_scratch: b200d_fuse_scratch_bytes(H,W) bytes. */
B200D_API size_t b200d_fuse_scratch_bytes(int H, int W);
B200D_API int b200d_mask_backproject(const float *d_depth, const float *d_conf, const uint8_t *d_bgr, int H, int W,
                           const double *Kinv64, const double *R64, const double *t64,
                           float min_conf, int mode, double *d_points, uint8_t *d_colors,
                           int64_t *d_count, void *d_scratch, void *stream);

/* Count-only pass of the same mask (no points written): d_count (1 x int64) receives the number of valid pixels, so a
 * caller can size one exact output buffer for many views with a single host synchronisation and then let
 * b200d_mask_backproject write every view into its slice. */
B200D_API int b200d_mask_count(const float *d_depth, const float *d_conf, int H, int W, float min_conf, int mode,
                               int64_t *d_count, void *d_scratch, void *stream);

/* Replaces `_prepare_images` (mvs_patchmatch.py:167-191, dense_stereo.py:156-176) for one image:
 * cv.resize(INTER_LINEAR) of the BGR uint8 image to (nh,nw), BGR2GRAY, /255 -- OpenCV's fixed-point arithmetic,
 * bit-exact.  d_xi/d_xa (nw / nw*2) and d_yi/d_ya (nh / nh*2) are the source-index and 11-bit coefficient tables
 * OpenCV builds per output column / row (b200dense/prepare.py builds them).  Outputs: d_color (nh,nw,3) u8 BGR,
 * d_gray (nh,nw) fp32 in [0,1]. */
B200D_API int b200d_prepare_image(const uint8_t *d_bgr, int H, int W, int nh, int nw, const int32_t *d_xi,
                                  const int16_t *d_xa, const int32_t *d_yi, const int16_t *d_ya, uint8_t *d_color,
                                  float *d_gray, void *stream);

/* Replaces the k-NN query of DenseStereoReconstructor._filter_outliers (dense_stereo.py:456-460):
 * d_mean_dist[i] = mean of the distances from point i to its k-1 nearest other points (f64, the values and the
 * summation order scikit-learn's KD-tree query + np.mean produce).  d_points (n,3) f64; 2 <= k <= 32, k <= n. */
B200D_API int b200d_knn_mean_dist(const double *d_points, int64_t n, int k, double *d_mean_dist, void *stream);

/* HOST function (no CUDA): replaces the per-point Python loop of save_ply (src/core/utils.py:8-37).  Writes the
 * same bytes: ASCII PLY header, then "%.6f %.6f %.6f %d %d %d" per vertex.  points (n,3) f64, colors (n,3) u8 RGB.
 * Vertices are formatted in parallel on the host threads (OpenMP) and written in order. */
B200D_API int b200d_write_ply_ascii(const char *path, const double *points, const uint8_t *colors, int64_t n);

/* EXTENSION (no counterpart in the reference, whose consistency check is photometric -- SURVEY F6): cross-view
 * geometric consistency.  For every reference pixel: back-project at d_depth_ref, project into each source,
 * read that source's depth map at the nearest pixel, back-project, return to the reference camera; a source is
 * consistent when the round trip lands within px_thresh pixels and rel_thresh relative depth.  d_count (H,W)
 * receives the number of consistent sources.  cams: K, R_ref/t_ref and the sources' R/t (Kinv unused). */
B200D_API int b200d_geo_consistency(const float *d_depth_ref, const float *const *d_depth_src, const b200d_cams *cams,
                                    int H, int W, float px_thresh, float rel_thresh, float *d_count, void *stream);

/* Diagnostic (synchronises): checks the branch-free IEEE sqrt / division helpers the NCC epilogue uses against the
 * compiler's __fsqrt_rn / __fdiv_rn -- the square root over all 2^32 bit patterns, the division over n_div
 * pseudo-random (numerator, divisor) pairs with any numerator and the divisors its two call sites produce.
 * out8 (HOST) = {sqrt mismatches among trusted results, trusted sqrt results, division mismatches, trusted divisions,
 * mismatches / count of the sampling kernel's shared-reciprocal division inside its range guard, mismatches / count of
 * `a / (W-1) - 1` under the relaxed (upper-bound-only) guard}; untrusted results are the ones the kernels redo with
 * the intrinsics. */
B200D_API int b200d_selftest_fast_math(uint64_t n_div, uint64_t *out8, void *stream);

/* Diagnostic (synchronises): the FP32 roofline denominator measured on the current device -- sustained packed-FMA
 * (FFMA2) rate in T lane-FMA/s (x2 = TFLOP/s), from a loop of independent fma.rn.f32x2 chains that exercises nothing
 * but the FMA pipe (csrc/diag_kernels.cuh; nominal 148 SMs x 128 lanes x clock). */
B200D_API int b200d_measure_fp32_peak(double *tlaneops, void *stream);

/* Optional per-launch timing for benchmarks.  Between _begin and _end every tile kernel launched by
 * this library is bracketed by CUDA events on its own stream (library-owned events, at most
 * max_records launches).  _end waits for the recorded launches, writes their durations (ms) and tags
 * (NCC kernels: mode*100 + patch with mode 0 stats, 1 cost, 2 confidence, 3 shift, 4 refine, 5 plane
 * sweep, 6 plane sweep with the NCC dump, 7 view propagation; +10000 for the B200D_FAST twins, +20000 for the bounded
 * plane sweep; sampling kernels: 2000 + 10*candidate + bounds (12000 + ... in B200D_FAST); init 3000; fusion 4000;
 * geometric consistency 5000; view-propagation proposals 5100) and releases the events.  Not meant to be left on in production. */
B200D_API int b200d_profile_begin(int max_records);
B200D_API int b200d_profile_end(float *ms_out, int32_t *tag_out, int cap, int *n_out);

#ifdef __cplusplus
}
#endif
#endif /* B200DENSE_H */
