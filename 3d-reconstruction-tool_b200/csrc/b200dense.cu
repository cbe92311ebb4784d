// b200dense.cu -- C-ABI entry points (include/b200dense.h) and kernel dispatch.
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false (see build.py).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "b200dense.h"
#include "fuse_kernels.cuh"
#include "pm_kernels.cuh"
#include "sweep_kernels.cuh"

using namespace b200d;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

int cuda_check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return B200D_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return B200D_ERR_CUDA;
}

#define CUDA_TRY(expr)                                        \
    do {                                                      \
        int _rc = cuda_check((expr), #expr);                  \
        if (_rc != B200D_OK) return _rc;                      \
    } while (0)

size_t align256(size_t n) { return (n + 255) / 256 * 256; }

// ---- optional per-launch timing (b200d_profile_begin/_end): CUDA events recorded on the launching
// stream around every tile kernel, so a benchmark can report the dominant kernel's own duration.
struct ProfRec { cudaEvent_t e0, e1; int tag; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
size_t g_prof_used = 0;
bool g_prof_on = false;

struct ProfScope {
    ProfRec *r = nullptr;
    cudaStream_t st;
    ProfScope(int tag, cudaStream_t s) : st(s) {
        if (!g_prof_on) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (g_prof_used < g_prof.size()) {
            r = &g_prof[g_prof_used++];
            r->tag = tag;
            cudaEventRecord(r->e0, st);
        }
    }
    ~ProfScope() {
        if (r) cudaEventRecord(r->e1, st);
    }
};

dim3 tile_grid(int H, int W) { return dim3((W + TW - 1) / TW, (H + TH - 1) / TH, 1); }

template <int K, int MODE>
int launch_pm(const PmArgs &a, cudaStream_t st) {
    static bool configured = false;          // per instantiation; attribute is idempotent
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(pm_eval_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)Geo<K>::smem_bytes));
        configured = true;
    }
    ProfScope prof(MODE * 100 + K, st);
    pm_eval_kernel<K, MODE><<<tile_grid(a.H, a.W), NT, Geo<K>::smem_bytes, st>>>(a);
    return cuda_check(cudaGetLastError(), "pm_eval_kernel launch");
}

template <int MODE>
int launch_pm_k(int patch, const PmArgs &a, cudaStream_t st) {
    switch (patch) {
        case 3: return launch_pm<3, MODE>(a, st);
        case 5: return launch_pm<5, MODE>(a, st);
        case 7: return launch_pm<7, MODE>(a, st);
        case 9: return launch_pm<9, MODE>(a, st);
        case 11: return launch_pm<11, MODE>(a, st);
        case 13: return launch_pm<13, MODE>(a, st);
        default: return fail(B200D_ERR_ARG, "unsupported patch size (odd 3..13)%s");
    }
}

template <int K>
int launch_sweep(const SweepArgs &a, cudaStream_t st) {
    constexpr size_t smem = 3 * (size_t)Geo<K>::RH * Geo<K>::RWP * sizeof(float) + (size_t)TH * TW;
    static bool configured = false;
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(plane_sweep_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    ProfScope prof(1000 + K, st);
    plane_sweep_kernel<K><<<tile_grid(a.H, a.W), NT, smem, st>>>(a);
    return cuda_check(cudaGetLastError(), "plane_sweep_kernel launch");
}

int check_common(int H, int W, int patch) {
    if (H < 2 || W < 2) return fail(B200D_ERR_ARG, "image must be at least 2x2%s");
    if (patch < 3 || patch > 13 || (patch % 2) == 0) return fail(B200D_ERR_ARG, "unsupported patch size (odd 3..13)%s");
    return B200D_OK;
}

__global__ void pm_init_kernel(float *depth, float *normal, float *best, size_t n, float log_span, float log_min,
                               const float *init_depth, const float *init_normal, uint64_t seed, uint32_t view_id) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 r = make_uint4(0, 0, 0, 0);
    if (!init_depth || !init_normal) r = philox4x32_10(make_uint4((uint32_t)i, 0xFFFFFFFFu, view_id, 0u), key);
    // depth = exp(rand * (ln dmax - ln dmin) + ln dmin)                     (mvs_patchmatch.py:268-272)
    depth[i] = init_depth ? init_depth[i] : expf(add(mul(u01(r.x), log_span), log_min));
    if (init_normal) {
        normal[3 * i] = init_normal[3 * i];
        normal[3 * i + 1] = init_normal[3 * i + 1];
        normal[3 * i + 2] = init_normal[3 * i + 2];
    } else {
        // normal = normalize([0.3 randn, 0.3 randn, -1])                     (:275-281)
        const uint4 r2 = philox4x32_10(make_uint4((uint32_t)i, 0xFFFFFFFFu, view_id, 1u), key);
        const float a = mul(gauss(r.y, r.z), 0.3f), b = mul(gauss(r2.x, r2.y), 0.3f);
        float nn = __fsqrt_rn(add(add(mul(a, a), mul(b, b)), 1.0f));
        nn = nn < 1e-12f ? 1e-12f : nn;
        normal[3 * i] = div_(a, nn);
        normal[3 * i + 1] = div_(b, nn);
        normal[3 * i + 2] = div_(-1.0f, nn);
    }
    best[i] = __int_as_float(0x7f800000);                                     // +inf (:284)
}

void fill_srcs(PmArgs &a, const float *const *d_src, int n) {
    for (int s = 0; s < B200D_MAX_SRC; ++s) a.src[s] = (s < n) ? d_src[s] : nullptr;
}

}  // namespace

extern "C" {

int b200d_abi_version(void) { return B200D_ABI_VERSION; }
const char *b200d_last_error(void) { return g_err; }

int b200d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200d_profile_begin(int max_records) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (max_records < 1 || max_records > (1 << 20)) return fail(B200D_ERR_ARG, "max_records out of range%s");
    for (auto &r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.assign((size_t)max_records, ProfRec{});
    for (auto &r : g_prof) {
        CUDA_TRY(cudaEventCreate(&r.e0));
        CUDA_TRY(cudaEventCreate(&r.e1));
    }
    g_prof_used = 0;
    g_prof_on = true;
    return B200D_OK;
}

int b200d_profile_end(float *ms_out, int32_t *tag_out, int cap, int *n_out) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = false;
    int n = 0, rc = B200D_OK;
    for (size_t i = 0; i < g_prof_used && rc == B200D_OK; ++i) {
        float ms = 0.f;
        rc = cuda_check(cudaEventSynchronize(g_prof[i].e1), "cudaEventSynchronize");
        if (rc == B200D_OK) rc = cuda_check(cudaEventElapsedTime(&ms, g_prof[i].e0, g_prof[i].e1), "cudaEventElapsedTime");
        if (rc == B200D_OK && n < cap && ms_out && tag_out) { ms_out[n] = ms; tag_out[n] = g_prof[i].tag; ++n; }
    }
    for (auto &r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_prof_used = 0;
    if (n_out) *n_out = n;
    return rc;
}

size_t b200d_pm_scratch_bytes(int H, int W) {
    const size_t n = align256((size_t)H * W * sizeof(float));
    return 12 * n;          // depth x2, best x2, mean1, var1 (1 plane each) + normal x2 (3 planes each)
}

size_t b200d_sweep_scratch_bytes(int H, int W) { return 2 * (size_t)H * W * sizeof(float); }

size_t b200d_fuse_scratch_bytes(int H, int W) {
    const size_t chunks = ((size_t)H * W + FUSE_CHUNK - 1) / FUSE_CHUNK;
    return align256(chunks * sizeof(uint32_t)) + align256(chunks * sizeof(int64_t));
}

int b200d_ref_stats(const float *d_ref, int H, int W, int patch, float *d_mean1, float *d_var1, void *stream) {
    int rc = check_common(H, W, patch);
    if (rc) return rc;
    if (!d_ref || !d_mean1 || !d_var1) return fail(B200D_ERR_ARG, "null pointer%s");
    PmArgs a;
    memset(&a, 0, sizeof(a));
    a.ref = d_ref; a.H = H; a.W = W; a.out = d_mean1; a.out2 = d_var1;
    return launch_pm_k<M_STATS>(patch, a, (cudaStream_t)stream);
}

int b200d_pm_cost(const float *d_ref, const float *d_mean1, const float *d_var1, const float *const *d_src,
                  const b200d_cams *cams, const float *d_depth, int H, int W, int patch, int mode, int arith,
                  float *d_out, void *stream) {
    int rc = check_common(H, W, patch);
    if (rc) return rc;
    if (!d_ref || !d_mean1 || !d_var1 || !d_src || !cams || !d_depth || !d_out) return fail(B200D_ERR_ARG, "null pointer%s");
    if (cams->n_src < 1 || cams->n_src > B200D_MAX_SRC) return fail(B200D_ERR_ARG, "n_src out of range%s");
    if (arith != B200D_PARITY) return fail(B200D_ERR_UNSUPPORTED, "fast arithmetic not built yet%s");
    PmArgs a;
    memset(&a, 0, sizeof(a));
    a.ref = d_ref; a.mean1 = d_mean1; a.var1 = d_var1; a.cams = *cams; a.H = H; a.W = W;
    fill_srcs(a, d_src, cams->n_src);
    a.depth_in = d_depth; a.out = d_out;
    if (mode == 0) return launch_pm_k<M_COST>(patch, a, (cudaStream_t)stream);
    if (mode == 1) return launch_pm_k<M_CONF>(patch, a, (cudaStream_t)stream);
    return fail(B200D_ERR_ARG, "mode must be 0 (cost) or 1 (confidence)%s");
}

int b200d_pm_view(const float *d_ref, const float *const *d_src, const b200d_cams *cams, int H, int W,
                  const b200d_pm_params *prm, const b200d_rng *rng, void *d_scratch, float *d_depth, float *d_normal,
                  float *d_conf, float *d_best, float *d_costs, void *stream) {
    if (!prm) return fail(B200D_ERR_ARG, "null params%s");
    int rc = check_common(H, W, prm->patch);
    if (rc) return rc;
    if (!d_ref || !d_src || !cams || !rng || !d_scratch || !d_depth || !d_normal || !d_conf)
        return fail(B200D_ERR_ARG, "null pointer%s");
    if (cams->n_src < 1 || cams->n_src > B200D_MAX_SRC) return fail(B200D_ERR_ARG, "n_src out of range%s");
    if (prm->iters < 0 || prm->samples < 0) return fail(B200D_ERR_ARG, "negative iteration count%s");
    if (prm->arith != B200D_PARITY) return fail(B200D_ERR_UNSUPPORTED, "fast arithmetic not built yet%s");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    const size_t plane = align256(n * sizeof(float));
    char *base = (char *)d_scratch;
    float *depth[2] = {(float *)(base), (float *)(base + plane)};
    float *best[2] = {(float *)(base + 2 * plane), (float *)(base + 3 * plane)};
    float *mean1 = (float *)(base + 4 * plane), *var1 = (float *)(base + 5 * plane);
    float *normal[2] = {(float *)(base + 6 * plane), (float *)(base + 9 * plane)};

    rc = b200d_ref_stats(d_ref, H, W, prm->patch, mean1, var1, stream);
    if (rc) return rc;

    const double ld0 = log(prm->depth_min), ld1 = log(prm->depth_max);
    pm_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(depth[0], normal[0], best[0], n, (float)(ld1 - ld0),
                                                                (float)ld0, rng->d_init_depth, rng->d_init_normal,
                                                                rng->seed, rng->view_id);
    CUDA_TRY(cudaGetLastError());

    PmArgs a;
    memset(&a, 0, sizeof(a));
    a.ref = d_ref; a.mean1 = mean1; a.var1 = var1; a.cams = *cams; a.H = H; a.W = W;
    fill_srcs(a, d_src, cams->n_src);
    a.depth_min = (float)prm->depth_min;
    a.depth_max = (float)prm->depth_max;
    a.seed = rng->seed;
    a.view_id = rng->view_id;

    int cd = 0, cn = 0, cb = 0, e = 0;
    for (int it = 0; it < prm->iters; ++it) {
        const int sgn = (it % 2 == 0) ? 1 : -1;       // slicing semantics of :429-442
        for (int o = 0; o < 2; ++o) {
            a.dy = (o == 0) ? sgn : 0;
            a.dx = (o == 0) ? 0 : sgn;
            a.depth_in = depth[cd]; a.normal_in = normal[cn]; a.best_in = best[cb];
            a.depth_out = depth[cd ^ 1]; a.normal_out = normal[cn ^ 1]; a.best_out = best[cb ^ 1];
            a.out = d_costs ? d_costs + (size_t)e * n : nullptr;
            a.eval_id = (uint32_t)e;
            rc = launch_pm_k<M_SHIFT>(prm->patch, a, st);
            if (rc) return rc;
            cd ^= 1; cn ^= 1; cb ^= 1; ++e;
        }
        a.dr = (float)((prm->depth_max - prm->depth_min) * pow(0.5, it));      // :466
        a.nr = (float)(0.5 * pow(0.5, it));                                    // :467
        for (int sm = 0; sm < prm->samples; ++sm) {
            const size_t k = (size_t)it * prm->samples + sm;
            a.rand_u = rng->d_rand_u ? rng->d_rand_u + k * n : nullptr;
            a.rand_n = rng->d_rand_n ? rng->d_rand_n + k * n * 3 : nullptr;
            a.depth_in = depth[cd]; a.depth_out = depth[cd ^ 1];
            a.normal_in = a.normal_out = normal[cn];                            // pixel-local: in place
            a.best_in = a.best_out = best[cb];
            a.out = d_costs ? d_costs + (size_t)e * n : nullptr;
            a.eval_id = (uint32_t)e;
            rc = launch_pm_k<M_REFINE>(prm->patch, a, st);
            if (rc) return rc;
            cd ^= 1; ++e;
        }
    }
    a.depth_in = depth[cd];
    a.out = d_conf;
    rc = launch_pm_k<M_CONF>(prm->patch, a, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_depth, depth[cd], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_normal, normal[cn], 3 * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (d_best) CUDA_TRY(cudaMemcpyAsync(d_best, best[cb], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return B200D_OK;
}

int b200d_plane_sweep(const float *d_ref, const float *const *d_nbr, const b200d_cams *cams, const float *d_depths,
                      int D, int H, int W, int patch, float thresh, int arith, float *d_depth_map, float *d_conf,
                      int32_t *d_idx, float *d_ncc, void *d_scratch, void *stream) {
    int rc = check_common(H, W, patch);
    if (rc) return rc;
    if (!d_ref || !d_nbr || !cams || !d_depths || !d_depth_map || !d_conf || !d_idx || !d_scratch) return fail(B200D_ERR_ARG, "null pointer%s");
    if (D < 1) return fail(B200D_ERR_ARG, "need at least one depth plane%s");
    if (cams->n_src < 1 || cams->n_src > B200D_MAX_SRC) return fail(B200D_ERR_ARG, "n_src out of range%s");
    if (arith != B200D_PARITY) return fail(B200D_ERR_UNSUPPORTED, "fast arithmetic not built yet%s");
    const size_t n = (size_t)H * W;
    float *stats = (float *)d_scratch;               // mean1 | var1 of the reference view
    rc = b200d_ref_stats(d_ref, H, W, patch, stats, stats + n, stream);
    if (rc == B200D_OK) {
        SweepArgs a;
        memset(&a, 0, sizeof(a));
        a.ref = d_ref; a.mean1 = stats; a.var1 = stats + n; a.cams = *cams; a.H = H; a.W = W; a.D = D;
        for (int s = 0; s < B200D_MAX_SRC; ++s) a.nbr[s] = (s < cams->n_src) ? d_nbr[s] : nullptr;
        a.depths = d_depths; a.thresh = thresh;
        a.depth_map = d_depth_map; a.conf = d_conf; a.idx = d_idx; a.ncc = d_ncc;
        cudaStream_t st = (cudaStream_t)stream;
        switch (patch) {
            case 3: rc = launch_sweep<3>(a, st); break;
            case 5: rc = launch_sweep<5>(a, st); break;
            case 7: rc = launch_sweep<7>(a, st); break;
            case 9: rc = launch_sweep<9>(a, st); break;
            case 11: rc = launch_sweep<11>(a, st); break;
            case 13: rc = launch_sweep<13>(a, st); break;
            default: rc = fail(B200D_ERR_ARG, "unsupported patch size (odd 3..13)%s");
        }
    }
    return rc;
}

int b200d_mask_backproject(const float *d_depth, const float *d_conf, const uint8_t *d_bgr, int H, int W,
                           const double *Kinv64, const double *R64, const double *t64, float min_conf, int mode,
                           double *d_points, uint8_t *d_colors, int64_t *d_count, void *d_scratch, void *stream) {
    if (H < 1 || W < 1) return fail(B200D_ERR_ARG, "empty image%s");
    if (!d_depth || !d_conf || !d_bgr || !Kinv64 || !R64 || !t64 || !d_points || !d_colors || !d_count || !d_scratch)
        return fail(B200D_ERR_ARG, "null pointer%s");
    if (mode != 0 && mode != 1) return fail(B200D_ERR_ARG, "mode must be 0 (PatchMatch) or 1 (plane sweep)%s");
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    a.depth = d_depth; a.conf = d_conf; a.bgr = d_bgr; a.H = H; a.W = W; a.min_conf = min_conf; a.mode = mode;
    memcpy(a.Kinv, Kinv64, sizeof(a.Kinv));
    memcpy(a.R, R64, sizeof(a.R));
    memcpy(a.t, t64, sizeof(a.t));
    a.points = d_points; a.colors = d_colors; a.count = d_count;
    const size_t chunks = ((size_t)H * W + FUSE_CHUNK - 1) / FUSE_CHUNK;
    a.n_chunks = (int)chunks;
    a.chunk_counts = (uint32_t *)d_scratch;
    a.chunk_offsets = (int64_t *)((char *)d_scratch + align256(chunks * sizeof(uint32_t)));
    cudaStream_t st = (cudaStream_t)stream;
    fuse_count_kernel<<<(unsigned)chunks, FUSE_NT, 0, st>>>(a);
    fuse_scan_kernel<<<1, 1024, 0, st>>>(a);
    fuse_write_kernel<<<(unsigned)chunks, FUSE_NT, 0, st>>>(a);
    return cuda_check(cudaGetLastError(), "fuse kernels launch");
}

}  // extern "C"
