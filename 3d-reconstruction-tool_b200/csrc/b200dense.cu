// b200dense.cu -- C-ABI entry points (include/b200dense.h) and kernel dispatch.
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false (see build.py).
//
// Every cost evaluation is two launches on the caller's stream:
//   sample_kernel  (sample_kernels.cuh)  candidate depth -> reproject -> bilinear sample, per pixel
//   ncc_kernel     (ncc_kernels.cuh)     TMA-fed tiles -> K x K window sums (FFMA2) -> NCC -> select
// The sampled images sit in an L2-sized scratch between the two.
#include <atomic>
#include <fcntl.h>
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "b200dense.h"
#include "diag_kernels.cuh"
#include "fast_kernels.cuh"
#include "fuse_kernels.cuh"
#include "geo_kernels.cuh"
#include "knn_kernels.cuh"
#include "ncc_kernels.cuh"
#include "prep_kernels.cuh"
#include "sample_kernels.cuh"

using namespace b200d;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
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
#define TRY(expr)                                             \
    do {                                                      \
        int _rc = (expr);                                     \
        if (_rc != B200D_OK) return _rc;                      \
    } while (0)

size_t align256(size_t n) { return (n + 255) / 256 * 256; }

// ---- optional per-launch timing (b200d_profile_begin/_end): CUDA events recorded on the launching
// stream around every kernel, so a benchmark can report each kernel's own duration.
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

// ---- scratch layout shared by all entry points (one slot = one view in flight)
struct Scratch {
    float *mean1, *var1, *cand, *best, *normal_alt, *rpad, *smap;
    uint8_t *valid, *allow;
    int Wp;
    size_t map_stride;           // floats between consecutive sampled maps
    static int pitch(int W) { return (W + 3 + 3) / 4 * 4; }     // room for the largest column offset (PAD <= 3)
    static size_t bytes(int H, int W) {
        const size_t n = (size_t)H * W, np = (size_t)H * pitch(W);
        return 4 * align256(n * 4) + align256(3 * n * 4) + 2 * align256(n) + align256(np * 4) +
               B200D_MAX_SRC * align256(np * 4);
    }
    Scratch(void *base, int H, int W) {
        const size_t n = (size_t)H * W, np = (size_t)H * pitch(W);
        char *p = (char *)(((uintptr_t)base + 255) & ~(uintptr_t)255);
        auto take = [&](size_t b) { char *q = p; p += align256(b); return q; };
        mean1 = (float *)take(n * 4);
        var1 = (float *)take(n * 4);
        cand = (float *)take(n * 4);
        best = (float *)take(n * 4);
        normal_alt = (float *)take(3 * n * 4);
        valid = (uint8_t *)take(n);
        allow = (uint8_t *)take(n);
        rpad = (float *)take(np * 4);
        smap = (float *)take(np * 4);
        Wp = pitch(W);
        map_stride = align256(np * 4) / 4;
    }
};

// ---- TMA descriptors (driver entry point resolved at run time: no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(CUtensorMap *m, const float *base, int W, int H, size_t pitch_floats, int box_w, int box_h) {
    // resolved once per process; C++11 guarantees the initialisation of a function-local static is thread-safe
    static const EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    if (!fn) {
        cudaGetLastError();
        return fail(B200D_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    }
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_floats * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (CUresult %d) W=%d H=%d pitch=%zu box=%dx%d", (int)r, W,
                 H, pitch_floats, box_w, box_h);
        return B200D_ERR_CUDA;
    }
    return B200D_OK;
}

// window (TMA box) geometry and column offset of a patch size
int box_dims(int patch, int *bw, int *bh, int *pad) {
    switch (patch) {
        case 3: *bw = Geo<3>::RWP; *bh = Geo<3>::RH; *pad = Geo<3>::PAD; return B200D_OK;
        case 5: *bw = Geo<5>::RWP; *bh = Geo<5>::RH; *pad = Geo<5>::PAD; return B200D_OK;
        case 7: *bw = Geo<7>::RWP; *bh = Geo<7>::RH; *pad = Geo<7>::PAD; return B200D_OK;
        case 9: *bw = Geo<9>::RWP; *bh = Geo<9>::RH; *pad = Geo<9>::PAD; return B200D_OK;
        case 11: *bw = Geo<11>::RWP; *bh = Geo<11>::RH; *pad = Geo<11>::PAD; return B200D_OK;
        case 13: *bw = Geo<13>::RWP; *bh = Geo<13>::RH; *pad = Geo<13>::PAD; return B200D_OK;
        default: return fail(B200D_ERR_ARG, "unsupported patch size (odd 3..13)");
    }
}

// Prepare the TMA sources of one view: a column-shifted copy of the reference image, zeroed left
// margins (the columns that stand for x < 0) and the tensor maps of the reference and of S sampled maps.
int prepare_tma(int patch, const float *d_ref, int H, int W, int S, const Scratch &sc, NccArgs &n, cudaStream_t st) {
    int bw, bh, pad;
    TRY(box_dims(patch, &bw, &bh, &pad));
    CUDA_TRY(cudaMemcpy2DAsync(sc.rpad + pad, (size_t)sc.Wp * 4, d_ref, (size_t)W * 4, (size_t)W * 4, H,
                               cudaMemcpyDeviceToDevice, st));
    if (pad) {
        CUDA_TRY(cudaMemset2DAsync(sc.rpad, (size_t)sc.Wp * 4, 0, (size_t)pad * 4, H, st));
        for (int s = 0; s < S; ++s)
            CUDA_TRY(cudaMemset2DAsync(sc.smap + s * sc.map_stride, (size_t)sc.Wp * 4, 0, (size_t)pad * 4, H, st));
    }
    TRY(make_tensor_map(&n.tm_r, sc.rpad, W + pad, H, sc.Wp, bw, bh));
    for (int s = 0; s < S; ++s) TRY(make_tensor_map(&n.tm_s[s], sc.smap + s * sc.map_stride, W + pad, H, sc.Wp, bw, bh));
    return B200D_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device (and per kernel): one bit per device ordinal, set once
template <typename Kern>
int ensure_smem(Kern kern, std::atomic<uint64_t> &configured, size_t bytes) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured.fetch_or(bit, std::memory_order_release);
    }
    return B200D_OK;
}

template <int K, int MODE>
int launch_ncc(const NccArgs &a, cudaStream_t st) {
    static std::atomic<uint64_t> configured{0};
    TRY(ensure_smem(ncc_kernel<K, MODE>, configured, Geo<K>::smem_bytes));
    ProfScope prof(MODE * 100 + K, st);
    const dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, 1);
    ncc_kernel<K, MODE><<<grid, NT, Geo<K>::smem_bytes, st>>>(a);
    return cuda_check(cudaGetLastError(), "ncc_kernel launch");
}

template <int MODE>
int launch_ncc_k(int patch, const NccArgs &a, cudaStream_t st) {
    switch (patch) {
        case 3: return launch_ncc<3, MODE>(a, st);
        case 5: return launch_ncc<5, MODE>(a, st);
        case 7: return launch_ncc<7, MODE>(a, st);
        case 9: return launch_ncc<9, MODE>(a, st);
        case 11: return launch_ncc<11, MODE>(a, st);
        case 13: return launch_ncc<13, MODE>(a, st);
        default: return fail(B200D_ERR_ARG, "unsupported patch size (odd 3..13)");
    }
}

// separable-sum kernels (fast_kernels.cuh): ARITH = A_FAST (B200D_FAST) or A_HYBRID (plane sweep, parity results)
template <int K, int MODE, int ARITH>
int launch_sep(const NccArgs &a, cudaStream_t st) {
    static std::atomic<uint64_t> configured{0};
    constexpr size_t smem = SepCfg<K, MODE, ARITH>::smem_bytes;
    TRY(ensure_smem(ncc_sep_kernel<K, MODE, ARITH>, configured, smem));
    ProfScope prof((ARITH == A_FAST ? 10000 : ARITH == A_HYBRID ? 20000 : 30000) + MODE * 100 + K, st);
    const dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, 1);
    ncc_sep_kernel<K, MODE, ARITH><<<grid, NT, smem, st>>>(a);
    return cuda_check(cudaGetLastError(), "ncc_sep_kernel launch");
}

template <int MODE, int ARITH>
int launch_sep_k(int patch, const NccArgs &a, cudaStream_t st) {
    switch (patch) {
        case 3: return launch_sep<3, MODE, ARITH>(a, st);
        case 5: return launch_sep<5, MODE, ARITH>(a, st);
        case 7: return launch_sep<7, MODE, ARITH>(a, st);
        case 9: return launch_sep<9, MODE, ARITH>(a, st);
        case 11: return launch_sep<11, MODE, ARITH>(a, st);
        case 13: return launch_sep<13, MODE, ARITH>(a, st);
        default: return fail(B200D_ERR_ARG, "unsupported patch size (odd 3..13)");
    }
}

template <int CAND, int BOUNDS>
int launch_sample(const SampleArgs &a, cudaStream_t st) {
    ProfScope prof(2000 + CAND * 10 + BOUNDS, st);
    const dim3 grid((a.W + SBX - 1) / SBX, (a.H + SBY - 1) / SBY, 1), block(SBX, SBY, 1);
    switch (a.cams.n_src) {                  // the usual source counts get a fully unrolled kernel
        case 4: sample_kernel<CAND, BOUNDS, 4><<<grid, block, 0, st>>>(a); break;
        case 5: sample_kernel<CAND, BOUNDS, 5><<<grid, block, 0, st>>>(a); break;
        case 6: sample_kernel<CAND, BOUNDS, 6><<<grid, block, 0, st>>>(a); break;
        default: sample_kernel<CAND, BOUNDS, 0><<<grid, block, 0, st>>>(a); break;
    }
    return cuda_check(cudaGetLastError(), "sample_kernel launch");
}

// ---- one cost evaluation = sampling + NCC, in the arithmetic the caller chose.
// B200D_FAST composes the reprojection per source on the host, in double:
//   q = R_s (R_r^T (Kinv [x,y,1] d - t_r)) + t_s = d (R_s R_r^T Kinv) [x,y,1] + (t_s - R_s R_r^T t_r)
struct Evaluator {
    int arith = B200D_PARITY, patch = 0;
    cudaStream_t st = nullptr;
    SampleArgs s;
    NccArgs a;
    FastCams fc;

    void compose_fast() {
        const b200d_cams &c = s.cams;
        fc.n_src = c.n_src;
        for (int i = 0; i < 6; ++i) fc.K[i] = c.K[i];
        for (int v = 0; v < c.n_src; ++v) {
            double A[9], M[9];                                 // A = R_s R_r^T ; M = A Kinv
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    double acc = 0;
                    for (int k = 0; k < 3; ++k) acc += (double)c.R_src[v][i * 3 + k] * (double)c.R_ref[j * 3 + k];
                    A[i * 3 + j] = acc;
                }
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) {
                    double acc = 0;
                    for (int k = 0; k < 3; ++k) acc += A[i * 3 + k] * (double)c.Kinv[k * 3 + j];
                    M[i * 3 + j] = acc;
                }
                double t = (double)c.t_src[v][i];
                for (int k = 0; k < 3; ++k) t -= A[i * 3 + k] * (double)c.t_ref[k];
                fc.c[v][i] = (float)t;
            }
            for (int i = 0; i < 9; ++i) fc.M[v][i] = (float)M[i];
        }
    }

    template <int CAND, int BOUNDS>
    int sample() {
        if (arith != B200D_FAST) return launch_sample<CAND, BOUNDS>(s, st);
        FastSampleArgs f;
        memset(&f, 0, sizeof(f));
        for (int i = 0; i < s.cams.n_src; ++i) f.src[i] = s.src[i];
        f.cams = fc;
        f.H = s.H; f.W = s.W; f.Wp = s.Wp; f.pad = s.pad; f.map_stride = s.map_stride;
        f.depth_in = s.depth_in; f.dy = s.dy; f.dx = s.dx;
        f.depth_min = s.depth_min; f.depth_max = s.depth_max; f.dr = s.dr; f.plane_depth = s.plane_depth;
        f.rand_u = s.rand_u; f.seed = s.seed; f.view_id = s.view_id; f.eval_id = s.eval_id; f.half = s.half;
        f.smap = s.smap; f.cand = s.cand; f.valid = s.valid;
        ProfScope prof(12000 + CAND * 10 + BOUNDS, st);
        const dim3 grid((f.W + SBX - 1) / SBX, (f.H + SBY - 1) / SBY, 1), block(SBX, SBY, 1);
        sample_fast_kernel<CAND, BOUNDS><<<grid, block, 0, st>>>(f);
        return cuda_check(cudaGetLastError(), "sample_fast_kernel launch");
    }

    template <int MODE>
    int ncc() {
        if (arith == B200D_FAST) return launch_sep_k<MODE, A_FAST>(patch, a, st);
        return launch_ncc_k<MODE>(patch, a, st);
    }
};

int check_common(int H, int W, int patch) {
    if (H < 2 || W < 2) return fail(B200D_ERR_ARG, "image must be at least 2x2");
    if (patch < 3 || patch > 13 || (patch % 2) == 0) return fail(B200D_ERR_ARG, "unsupported patch size (odd 3..13)");
    return B200D_OK;
}

int check_cams(const b200d_cams *cams) {
    if (!cams) return fail(B200D_ERR_ARG, "null pointer");
    if (cams->n_src < 1 || cams->n_src > B200D_MAX_SRC) return fail(B200D_ERR_ARG, "n_src out of range");
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
        float g0, g1;
        gauss2(r.y, r.z, g0, g1);
        const float a = mul(g0, 0.3f), b = mul(g1, 0.3f);
        float nn = __fsqrt_rn(add(add(mul(a, a), mul(b, b)), 1.0f));
        nn = nn < 1e-12f ? 1e-12f : nn;
        normal[3 * i] = div_(a, nn);
        normal[3 * i + 1] = div_(b, nn);
        normal[3 * i + 2] = div_(-1.0f, nn);
    }
    best[i] = __int_as_float(0x7f800000);                                     // +inf (:284)
}

void fill_sample(SampleArgs &s, const float *const *d_src, const b200d_cams *cams, int H, int W, int patch,
                 const Scratch &sc) {
    memset(&s, 0, sizeof(s));
    for (int i = 0; i < cams->n_src; ++i) s.src[i] = d_src[i];
    s.cams = *cams;
    fill_pair_cams(s.pc, s.cams);
    s.H = H; s.W = W; s.Wp = sc.Wp; s.map_stride = sc.map_stride;
    s.half = patch / 2;
    s.pad = (patch / 2) % 4;
    s.smap = sc.smap; s.cand = sc.cand; s.valid = sc.valid;
}

// mean1 / var1: the "sampled image" of the statistics pass is the reference itself.
int stats_impl(const NccArgs &prepared, int H, int W, int patch, float *mean1, float *var1, cudaStream_t st) {
    NccArgs n;
    memset(&n, 0, sizeof(n));
    n.tm_r = prepared.tm_r;
    n.tm_s[0] = prepared.tm_r;
    n.n_src = 1; n.H = H; n.W = W; n.out = mean1; n.out2 = var1;
    return launch_ncc_k<N_STATS>(patch, n, st);
}

int check_arith(int arith) {
    if (arith != B200D_PARITY && arith != B200D_FAST && arith != B200D_PARITY_CHAINS && arith != B200D_PARITY_CHECK)
        return fail(B200D_ERR_ARG, "arith must be B200D_PARITY, B200D_FAST, B200D_PARITY_CHAINS or B200D_PARITY_CHECK");
    return B200D_OK;
}

// common set-up of an evaluation context: TMA sources, reference statistics, sampling / NCC argument blocks
int setup_eval(Evaluator &ev, const float *d_ref, const float *const *d_src, const b200d_cams *cams, int H, int W, int patch,
               int arith, const Scratch &sc, cudaStream_t st) {
    ev.arith = (arith == B200D_FAST) ? B200D_FAST : B200D_PARITY;
    ev.patch = patch;
    ev.st = st;
    memset(&ev.a, 0, sizeof(ev.a));
    TRY(prepare_tma(patch, d_ref, H, W, cams->n_src, sc, ev.a, st));
    TRY(stats_impl(ev.a, H, W, patch, sc.mean1, sc.var1, st));
    fill_sample(ev.s, d_src, cams, H, W, patch, sc);
    if (ev.arith == B200D_FAST) ev.compose_fast();
    NccArgs &a = ev.a;
    a.n_src = cams->n_src; a.H = H; a.W = W; a.mean1 = sc.mean1; a.var1 = sc.var1; a.valid = sc.valid; a.cand = sc.cand;
    a.color = -1;
    a.smap = sc.smap; a.rpad = sc.rpad; a.map_stride = sc.map_stride; a.Wp = sc.Wp; a.pad = (patch / 2) % 4;
    return B200D_OK;
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
    if (max_records < 1 || max_records > (1 << 20)) return fail(B200D_ERR_ARG, "max_records out of range");
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

size_t b200d_scratch_bytes(int H, int W) {
    if (H < 1 || W < 1) return 0;
    return Scratch::bytes(H, W) + 256;
}

size_t b200d_fuse_scratch_bytes(int H, int W) {
    const size_t chunks = ((size_t)H * W + FUSE_CHUNK - 1) / FUSE_CHUNK;
    return align256(chunks * sizeof(uint32_t)) + align256(chunks * sizeof(int64_t));
}

int b200d_ref_stats(const float *d_ref, int H, int W, int patch, float *d_mean1, float *d_var1, void *d_scratch,
                    void *stream) {
    TRY(check_common(H, W, patch));
    if (!d_ref || !d_mean1 || !d_var1 || !d_scratch) return fail(B200D_ERR_ARG, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(d_scratch, H, W);
    NccArgs n;
    memset(&n, 0, sizeof(n));
    TRY(prepare_tma(patch, d_ref, H, W, 0, sc, n, st));
    return stats_impl(n, H, W, patch, d_mean1, d_var1, st);
}

int b200d_pm_cost(const float *d_ref, const float *const *d_src, const b200d_cams *cams, const float *d_depth, int H,
                  int W, int patch, int mode, int arith, float *d_out, void *d_scratch, void *stream) {
    TRY(check_common(H, W, patch));
    TRY(check_cams(cams));
    TRY(check_arith(arith));
    if (!d_ref || !d_src || !d_depth || !d_out || !d_scratch) return fail(B200D_ERR_ARG, "null pointer");
    if (mode != 0 && mode != 1) return fail(B200D_ERR_ARG, "mode must be 0 (cost) or 1 (confidence)");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(d_scratch, H, W);
    Evaluator ev;
    TRY(setup_eval(ev, d_ref, d_src, cams, H, W, patch, arith, sc, st));
    ev.s.depth_in = d_depth;
    ev.a.out = d_out;
    if (mode == 0) {
        TRY((ev.sample<C_MAP, B_PATCH>()));
        return ev.ncc<N_COST>();
    }
    TRY((ev.sample<C_MAP, B_IMAGE>()));
    return ev.ncc<N_CONF>();
}

int b200d_pm_view(const float *d_ref, const float *const *d_src, const b200d_cams *cams, int H, int W,
                  const b200d_pm_params *prm, const b200d_rng *rng, void *d_scratch, float *d_depth, float *d_normal,
                  float *d_conf, float *d_best, float *d_costs, void *stream) {
    if (!prm) return fail(B200D_ERR_ARG, "null params");
    TRY(check_common(H, W, prm->patch));
    TRY(check_cams(cams));
    TRY(check_arith(prm->arith));
    if (!d_ref || !d_src || !rng || !d_scratch || !d_depth || !d_normal) return fail(B200D_ERR_ARG, "null pointer");
    if (prm->iters < 0 || prm->samples < 0) return fail(B200D_ERR_ARG, "negative iteration count");
    if (!(prm->depth_min > 0.0) || !std::isfinite(prm->depth_min) || !std::isfinite(prm->depth_max) ||
        !(prm->depth_max > prm->depth_min))
        return fail(B200D_ERR_ARG, "need finite 0 < depth_min < depth_max (the initial depths are log-uniform)");
    if (prm->schedule != B200D_SCHEDULE_REFERENCE && prm->schedule != B200D_SCHEDULE_REDBLACK)
        return fail(B200D_ERR_ARG, "unknown propagation schedule");
    // staged execution (view propagation interleaves the views between iterations): iterations [it_begin, it_end)
    const int it0 = prm->it_begin, it1 = prm->it_end > 0 ? prm->it_end : prm->iters;
    const bool do_init = !(prm->stage & B200D_STAGE_NO_INIT), do_conf = !(prm->stage & B200D_STAGE_NO_CONF);
    if (it0 < 0 || it0 > it1 || it1 > prm->iters) return fail(B200D_ERR_ARG, "need 0 <= it_begin <= it_end <= iters");
    const bool staged = !do_init || !do_conf || it0 != 0 || it1 != prm->iters;
    if (staged && !d_best) return fail(B200D_ERR_ARG, "staged execution keeps the best cost in the caller's d_best");
    if (do_conf && !d_conf) return fail(B200D_ERR_ARG, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    const int patch = prm->patch;
    Scratch sc(d_scratch, H, W);
    // State lives in the caller's output buffers: depth and best are updated in place (only winners
    // write).  Normals ping-pong across the two shifted evaluations of an iteration because a shift
    // reads the neighbour's normal; an even number of flips per iteration -> the state is in d_normal
    // at every iteration boundary.
    float *depth = d_depth, *best = d_best ? d_best : sc.best;
    float *normal[2] = {d_normal, sc.normal_alt};

    Evaluator ev;
    TRY(setup_eval(ev, d_ref, d_src, cams, H, W, patch, prm->arith, sc, st));
    SampleArgs &s = ev.s;
    NccArgs &a = ev.a;

    if (do_init) {
        const double ld0 = log(prm->depth_min), ld1 = log(prm->depth_max);
        ProfScope prof(3000, st);
        pm_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(depth, normal[0], best, n, (float)(ld1 - ld0), (float)ld0,
                                                                    rng->d_init_depth, rng->d_init_normal, rng->seed,
                                                                    rng->view_id);
        CUDA_TRY(cudaGetLastError());
    }

    s.depth_in = depth;
    s.depth_min = (float)prm->depth_min;
    s.depth_max = (float)prm->depth_max;
    s.seed = rng->seed;
    s.view_id = rng->view_id;
    a.depth_in = a.depth_out = depth;
    a.best_in = a.best_out = best;
    a.seed = rng->seed;
    a.view_id = rng->view_id;

    const int per_it = (prm->schedule == B200D_SCHEDULE_REDBLACK ? 8 : 2) + prm->samples;
    int cn = 0, e = it0 * per_it;
    // one shifted-neighbour evaluation: candidate = hypothesis of the pixel at (+dy,+dx); `color` restricts who may adopt it
    auto shift_eval = [&](int dy, int dx, int color) -> int {
        s.dy = a.dy = dy;
        s.dx = a.dx = dx;
        a.color = color;
        s.eval_id = a.eval_id = (uint32_t)e;
        TRY((ev.sample<C_SHIFT, B_PATCH>()));
        a.normal_in = normal[cn]; a.normal_out = normal[cn ^ 1];
        a.out = d_costs ? d_costs + (size_t)e * n : nullptr;
        TRY(ev.ncc<N_SHIFT>());
        cn ^= 1; ++e;
        return B200D_OK;
    };
    for (int it = it0; it < it1; ++it) {
        if (prm->schedule == B200D_SCHEDULE_REDBLACK) {
            static const int dirs[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
            for (int color = 0; color < 2; ++color)
                for (int d = 0; d < 4; ++d) TRY(shift_eval(dirs[d][0], dirs[d][1], color));
        } else {
            const int sgn = (it % 2 == 0) ? 1 : -1;       // slicing semantics of :429-442
            TRY(shift_eval(sgn, 0, -1));
            TRY(shift_eval(0, sgn, -1));
        }
        a.color = -1;
        s.dr = (float)((prm->depth_max - prm->depth_min) * pow(0.5, it));      // :466
        a.nr = (float)(0.5 * pow(0.5, it));                                    // :467
        for (int sm = 0; sm < prm->samples; ++sm) {
            const size_t k = (size_t)it * prm->samples + sm;
            s.rand_u = rng->d_rand_u ? rng->d_rand_u + k * n : nullptr;
            a.rand_n = rng->d_rand_n ? rng->d_rand_n + k * n * 3 : nullptr;
            s.eval_id = a.eval_id = (uint32_t)e;
            TRY((ev.sample<C_REFINE, B_PATCH>()));
            a.normal_in = a.normal_out = normal[cn];                            // pixel-local: in place
            a.out = d_costs ? d_costs + (size_t)e * n : nullptr;
            TRY(ev.ncc<N_REFINE>());
            ++e;
        }
    }
    if (cn != 0) CUDA_TRY(cudaMemcpyAsync(d_normal, normal[cn], 3 * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (!do_conf) return B200D_OK;
    TRY((ev.sample<C_MAP, B_IMAGE>()));
    a.out = d_conf;
    return ev.ncc<N_CONF>();
}

int b200d_pm_viewprop(const float *d_ref, const float *const *d_src, const b200d_cams *cams, int H, int W,
                      const b200d_pm_params *prm, const float *const *d_src_depth, const float *const *d_src_normal,
                      void *d_scratch, float *d_depth, float *d_normal, float *d_best, float *d_costs, void *stream) {
    if (!prm) return fail(B200D_ERR_ARG, "null params");
    TRY(check_common(H, W, prm->patch));
    TRY(check_cams(cams));
    TRY(check_arith(prm->arith));
    if (!d_ref || !d_src || !d_src_depth || !d_src_normal || !d_scratch || !d_depth || !d_normal || !d_best)
        return fail(B200D_ERR_ARG, "null pointer");
    if (!(prm->depth_min > 0.0) || !(prm->depth_max > prm->depth_min)) return fail(B200D_ERR_ARG, "bad depth range");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    Scratch sc(d_scratch, H, W);
    Evaluator ev;
    TRY(setup_eval(ev, d_ref, d_src, cams, H, W, prm->patch, prm->arith, sc, st));
    SampleArgs &s = ev.s;
    NccArgs &a = ev.a;
    s.depth_in = sc.cand;                     // the candidate map the proposal kernel writes
    s.cand = nullptr;                         // ... which the sampling kernel therefore must not rewrite
    a.depth_in = a.depth_out = d_depth;
    a.best_in = a.best_out = d_best;
    a.normal_in = sc.normal_alt;              // candidate normals
    a.normal_out = d_normal;
    a.allow = sc.allow;
    for (int v = 0; v < cams->n_src; ++v) {
        if (!d_src_depth[v] || !d_src_normal[v]) return fail(B200D_ERR_ARG, "null pointer");
        VpropArgs vp;
        memset(&vp, 0, sizeof(vp));
        vp.depth = d_depth; vp.normal = d_normal; vp.depth_src = d_src_depth[v]; vp.normal_src = d_src_normal[v];
        vp.cams = *cams; vp.s = v; vp.H = H; vp.W = W;
        vp.depth_min = (float)prm->depth_min; vp.depth_max = (float)prm->depth_max;
        vp.cand_depth = sc.cand; vp.cand_normal = sc.normal_alt; vp.allow = sc.allow;
        {
            ProfScope prof(5100, st);
            viewprop_cand_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, 1), dim3(32, 8, 1), 0, st>>>(vp);
        }
        CUDA_TRY(cudaGetLastError());
        TRY((ev.sample<C_MAP, B_PATCH>()));
        a.out = d_costs ? d_costs + (size_t)v * n : nullptr;
        TRY(ev.ncc<N_VPROP>());
    }
    return B200D_OK;
}

int b200d_plane_sweep(const float *d_ref, const float *const *d_nbr, const b200d_cams *cams, const float *depths_host,
                      int D, int H, int W, int patch, float thresh, int arith, float *d_depth_map, float *d_conf,
                      int32_t *d_idx, float *d_ncc, void *d_scratch, void *stream) {
    TRY(check_common(H, W, patch));
    TRY(check_cams(cams));
    TRY(check_arith(arith));
    if (!d_ref || !d_nbr || !depths_host || !d_depth_map || !d_conf || !d_idx || !d_scratch)
        return fail(B200D_ERR_ARG, "null pointer");
    if (D < 1) return fail(B200D_ERR_ARG, "need at least one depth plane");
    if ((size_t)D > (size_t)H * W) return fail(B200D_ERR_ARG, "more planes than pixels");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(d_scratch, H, W);
    Evaluator ev;
    TRY(setup_eval(ev, d_ref, d_nbr, cams, H, W, patch, arith, sc, st));
    SampleArgs &s = ev.s;
    NccArgs &a = ev.a;
    // plane depths: kernel argument per plane (host) and a device copy for the final gather; the
    // candidate map of the scratch is unused by the sweep and holds the D floats
    float *d_depths = sc.cand;
    CUDA_TRY(cudaMemcpyAsync(d_depths, depths_host, (size_t)D * sizeof(float), cudaMemcpyHostToDevice, st));
    s.cand = nullptr;
    a.cand = nullptr;
    a.thresh = thresh; a.n_planes = D; a.depths = d_depths;
    if (thresh > 1e-3f && thresh < 1e3f) {                    // quick comparison bounds (sweep_over, sweep_decide_bounded)
        // margin 1e-5 on the threshold: the computed ncc = fl(cov / fl(sqrt(x))) is within (1 +- 4u) of cov / sqrt(x)
        // and the squared comparison adds 2u, so any margin above ~3e-7 makes the quick decision imply the exact one
        a.t_lo = (float)(((1.0 - 1e-5) * thresh) * ((1.0 - 1e-5) * thresh));
        a.t_hi = (float)(((1.0 + 1e-5) * thresh) * ((1.0 + 1e-5) * thresh));
    }
    a.conf = d_conf; a.idx = d_idx; a.depth_map = d_depth_map; a.ncc = d_ncc;
    // PARITY: separable sums + rigorous bound decide `ncc > thresh`, undecided warps redo the exact chains (bit-exact
    // counts); PARITY_CHAINS: the exact chains everywhere (A/B reference of the former); the NCC dump needs the values
    // themselves, hence the chains.
    const bool hybrid = (arith == B200D_PARITY || arith == B200D_PARITY_CHECK) && !d_ncc && a.t_lo > 0.0f;
    for (int d = 0; d < D; ++d) {
        s.plane_depth = depths_host[d];
        a.plane = d;
        TRY((ev.sample<C_PLANE, B_DEPTH>()));
        if (arith == B200D_FAST) TRY((launch_sep_k<N_SWEEP, A_FAST>(patch, a, st)));
        else if (hybrid && arith == B200D_PARITY_CHECK) TRY((launch_sep_k<N_SWEEP, A_CHECK>(patch, a, st)));
        else if (hybrid) TRY((launch_sep_k<N_SWEEP, A_HYBRID>(patch, a, st)));
        else if (d_ncc) TRY(launch_ncc_k<N_SWEEP_DUMP>(patch, a, st));
        else TRY(launch_ncc_k<N_SWEEP>(patch, a, st));
    }
    return B200D_OK;
}

int b200d_sweep_bound_check(float *ratios3, int reset) {
    unsigned v[3] = {0, 0, 0};
    if (ratios3) {
        CUDA_TRY(cudaMemcpyFromSymbol(v, g_bound_ratio, sizeof(v)));
        memcpy(ratios3, v, sizeof(v));
    }
    if (reset) {
        v[0] = v[1] = v[2] = 0;
        CUDA_TRY(cudaMemcpyToSymbol(g_bound_ratio, v, sizeof(v)));
    }
    return B200D_OK;
}

int b200d_sweep_fallbacks(uint64_t *count, int reset) {
    unsigned long long v = 0;
    if (count) {
        CUDA_TRY(cudaMemcpyFromSymbol(&v, g_sweep_fallbacks, sizeof(v)));
        *count = (uint64_t)v;
    }
    if (reset) {
        v = 0;
        CUDA_TRY(cudaMemcpyToSymbol(g_sweep_fallbacks, &v, sizeof(v)));
    }
    return B200D_OK;
}

int b200d_geo_consistency(const float *d_depth_ref, const float *const *d_depth_src, const b200d_cams *cams, int H, int W,
                           float px_thresh, float rel_thresh, float *d_count, void *stream) {
    TRY(check_cams(cams));
    if (H < 1 || W < 1) return fail(B200D_ERR_ARG, "empty image");
    if (!d_depth_ref || !d_depth_src || !d_count) return fail(B200D_ERR_ARG, "null pointer");
    GeoArgs a;
    memset(&a, 0, sizeof(a));
    a.depth_ref = d_depth_ref;
    for (int s = 0; s < cams->n_src; ++s) {
        if (!d_depth_src[s]) return fail(B200D_ERR_ARG, "null pointer");
        a.depth_src[s] = d_depth_src[s];
    }
    a.cams = *cams; a.H = H; a.W = W; a.px_thresh = px_thresh; a.rel_thresh = rel_thresh; a.out = d_count;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(5000, st);
    geo_consistency_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, 1), dim3(32, 8, 1), 0, st>>>(a);
    return cuda_check(cudaGetLastError(), "geo_consistency_kernel launch");
}

int b200d_knn_mean_dist(const double *d_points, int64_t n, int k, double *d_mean_dist, void *stream) {
    if (!d_points || !d_mean_dist) return fail(B200D_ERR_ARG, "null pointer");
    if (n < 1) return fail(B200D_ERR_ARG, "no points");
    if (k < 2 || k > KNN_KMAX || k > n) return fail(B200D_ERR_ARG, "k must be in 2..32 and <= n");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(7000, st);
    const int64_t per_block = (int64_t)KNN_NT * KNN_Q;
    knn_mean_dist_kernel<<<(unsigned)((n + per_block - 1) / per_block), KNN_NT, 0, st>>>(d_points, n, k, d_mean_dist);
    return cuda_check(cudaGetLastError(), "knn_mean_dist_kernel launch");
}

// ---- self-test of the branch-free IEEE helpers (device_math.cuh): every trusted result must equal the intrinsic's
namespace {
__global__ void selftest_sqrt_kernel(unsigned long long *bad, unsigned long long *trusted_n) {
    unsigned long long nb = 0, nt = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)i);
        bool t;
        const float f = sqrt_fast(x, t);
        if (t) {
            ++nt;
            if (__float_as_uint(f) != __float_as_uint(__fsqrt_rn(x))) ++nb;
        }
    }
    if (nb) atomicAdd(bad, nb);
    atomicAdd(trusted_n, nt);
}
__global__ void selftest_div_kernel(unsigned long long *bad, unsigned long long *trusted_n, unsigned long long n) {
    unsigned long long nb = 0, nt = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 7u, 9u), make_uint2(1u, 2u));
        const float a = __uint_as_float(r.x);                                   // any bit pattern
        // divisors of the two call sites: sqrt(..) + 1e-8 in [1e-8, 4) and count + 1e-8 in {1e-8, 1, ..., 8}
        const float b = (r.w & 1u) ? 1e-8f + u01(r.y) * ((r.w & 2u) ? 4.0f : 1e-4f) : (float)(r.z % 9u) + 1e-8f;
        bool t;
        const float f = div_fast(a, b, t);
        if (t) {
            ++nt;
            if (__float_as_uint(f) != __float_as_uint(__fdiv_rn(a, b))) ++nb;
        }
    }
    if (nb) atomicAdd(bad, nb);
    atomicAdd(trusted_n, nt);
}
// the sampling kernel's shared-reciprocal divisions: operands inside the range guard -> same bits as __fdiv_rn; and the
// relaxed case of the constant divisors W-1 / H-1: a numerator below the guard (incl. zero and denormals) may give an
// inexact quotient, but `quotient - 1` -- all the kernel uses -- must not change
__global__ void selftest_sample_div_kernel(unsigned long long *out, unsigned long long n) {
    unsigned long long nb = 0, nt = 0, rb = 0, rt = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 11u, 13u), make_uint2(3u, 4u));
        const float a = __uint_as_float(r.x), b = __uint_as_float(r.y);
        if (in_range(a) && in_range(b)) {
            ++nt;
            if (__float_as_uint(div_rcp(a, b, rcp_refined(b))) != __float_as_uint(__fdiv_rn(a, b))) ++nb;
        }
        const float wm1 = (float)(1u + (r.z & 0xffffffu));                      // W-1 in [1, 2^24]
        const float rw = rcp_refined(wm1);
        if (fabsf(a) < 0x1p40f) {                                               // the relaxed guard (upper bound only)
            ++rt;
            if (__float_as_uint(sub(div_rcp(a, wm1, rw), 1.0f)) != __float_as_uint(sub(__fdiv_rn(a, wm1), 1.0f))) ++rb;
        }
    }
    atomicAdd(out + 0, nb); atomicAdd(out + 1, nt); atomicAdd(out + 2, rb); atomicAdd(out + 3, rt);
}
}  // namespace

int b200d_selftest_fast_math(uint64_t n_div, uint64_t *out8, void *stream) {
    if (!out8) return fail(B200D_ERR_ARG, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 8 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 8 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) {
        cudaFree(d);
        return cuda_check(e, "cudaMemsetAsync");
    }
    selftest_sqrt_kernel<<<148 * 8, 256, 0, st>>>(d + 0, d + 1);
    selftest_div_kernel<<<148 * 8, 256, 0, st>>>(d + 2, d + 3, (unsigned long long)n_div);
    selftest_sample_div_kernel<<<148 * 8, 256, 0, st>>>(d + 4, (unsigned long long)n_div);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out8, d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    return cuda_check(e, "fast-math self-test");
}

int b200d_measure_fp32_peak(double *tlaneops, void *stream) {
    if (!tlaneops) return fail(B200D_ERR_ARG, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 8, block = 256, iters = 20000;
    float *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)grid * block * sizeof(float)));
    cudaError_t e = cudaMemsetAsync(d, 0, (size_t)grid * block * sizeof(float), st);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 3 && e == cudaSuccess; ++rep) {                 // first repetition warms up
        cudaEventRecord(e0, st);
        fma_peak_kernel<<<grid, block, 0, st>>>(d, 0.5f, iters);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess && rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_check(e, "fp32 peak measurement");
    *tlaneops = (double)grid * block * iters * 16.0 / (best * 1e-3) / 1e12;
    return B200D_OK;
}

int b200d_prepare_image(const uint8_t *d_bgr, int H, int W, int nh, int nw, const int32_t *d_xi, const int16_t *d_xa,
                        const int32_t *d_yi, const int16_t *d_ya, uint8_t *d_color, float *d_gray, void *stream) {
    if (H < 1 || W < 1 || nh < 1 || nw < 1) return fail(B200D_ERR_ARG, "empty image");
    if (!d_bgr || !d_xi || !d_xa || !d_yi || !d_ya || !d_color || !d_gray) return fail(B200D_ERR_ARG, "null pointer");
    PrepArgs a;
    a.src = d_bgr; a.H = H; a.W = W; a.nh = nh; a.nw = nw; a.xi = d_xi; a.yi = d_yi; a.xa = d_xa; a.ya = d_ya;
    a.color = d_color; a.gray = d_gray;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(6000, st);
    prepare_image_kernel<<<dim3((nw + 31) / 32, (nh + 7) / 8, 1), dim3(32, 8, 1), 0, st>>>(a);
    return cuda_check(cudaGetLastError(), "prepare_image_kernel launch");
}

// Host-side: ASCII PLY in the reference's exact format (src/core/utils.py:8-37): "%.6f %.6f %.6f %d %d %d\n".
// The reference formats one point per Python iteration.  Here blocks of points are formatted in parallel (OpenMP:
// every thread formats a contiguous run of vertices into its own buffer) and every thread then writes its run at the
// file offset the run lengths imply (pwrite), so the file is byte-identical to the sequential writer's.
//
// "%.6f" of a double, exactly as the C library / Python print it: the decimal expansion of the binary value, rounded
// to six places, ties to even.  x = m 2^e with a 53-bit m, so for e < 0 the scaled value m 10^6 / 2^-e is a 128-bit
// integer division by a power of two: quotient, remainder against one half, done -- no floating point involved, ~10x
// faster than snprintf.  Returns NULL for what it does not cover (inf, NaN, |x| >= 2^52, results beyond 64 bits);
// the caller then uses snprintf for that value.
static inline char *fmt_u64(char *p, uint64_t v) {
    char tmp[20];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
static inline char *fmt_fixed6(char *p, double x) {
    uint64_t bits;
    memcpy(&bits, &x, sizeof(bits));
    const int ex = (int)((bits >> 52) & 0x7ff);
    uint64_t n = 0;                                       // round_half_even(|x| * 10^6)
    if (ex == 0x7ff) return nullptr;                      // inf / NaN
    if (ex != 0) {                                        // zero and subnormals print as 0.000000
        const uint64_t man = (bits & ((1ull << 52) - 1)) | (1ull << 52);
        const int sh = 1075 - ex;                         // |x| = man / 2^sh
        if (sh <= 0) return nullptr;                      // |x| >= 2^52
        if (sh < 128) {                                   // otherwise |x| 10^6 < 2^73 / 2^128: rounds to 0
            const unsigned __int128 P = (unsigned __int128)man * 1000000u;
            unsigned __int128 q = P >> sh;
            const unsigned __int128 rem = P - (q << sh), half = (unsigned __int128)1 << (sh - 1);
            if (rem > half || (rem == half && ((uint64_t)q & 1u))) ++q;
            if (q >> 64) return nullptr;
            n = (uint64_t)q;
        }
    }
    if (bits >> 63) *p++ = '-';
    p = fmt_u64(p, n / 1000000u);
    *p++ = '.';
    uint32_t fr = (uint32_t)(n % 1000000u);
    for (int i = 5; i >= 0; --i) { p[i] = (char)('0' + fr % 10); fr /= 10; }
    return p + 6;
}
// one coordinate + the separator that follows it; falls back to snprintf for the values fmt_fixed6 declines
static inline size_t fmt_coord(std::vector<char> &buf, size_t pos, double x) {
    if (buf.size() < pos + 64) buf.resize(buf.size() * 2 + (1 << 16));
    char *p = fmt_fixed6(buf.data() + pos, x);
    if (!p) {
        int w = snprintf(buf.data() + pos, buf.size() - pos, "%.6f", x);
        if (w < 0) w = 0;
        if ((size_t)w + 2 >= buf.size() - pos) {          // astronomically large coordinate: grow and redo
            buf.resize(buf.size() * 2 + (size_t)w + 1024);
            w = snprintf(buf.data() + pos, buf.size() - pos, "%.6f", x);
        }
        p = buf.data() + pos + w;
    }
    *p++ = ' ';
    return (size_t)(p - buf.data());
}

static bool write_all_at(int fd, const char *data, size_t len, off_t at) {
    while (len) {
        const ssize_t w = pwrite(fd, data, len, at);
        if (w <= 0) return false;
        data += w; len -= (size_t)w; at += w;
    }
    return true;
}

int b200d_write_ply_ascii(const char *path, const double *points, const uint8_t *colors, int64_t n) {
    if (!path || n < 0 || (n > 0 && (!points || !colors))) return fail(B200D_ERR_ARG, "bad argument");
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(B200D_ERR_ARG, "cannot open output file");
    char header[256];
    const int hl = snprintf(header, sizeof(header),
                            "ply\nformat ascii 1.0\nelement vertex %lld\nproperty float x\nproperty float y\nproperty float z\n"
                            "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n", (long long)n);
    bool bad = !write_all_at(fd, header, (size_t)hl, 0);
    off_t filepos = hl;
    // Threads: the cores this process may run on (affinity mask), not OMP_NUM_THREADS -- launchers such as torchrun
    // export OMP_NUM_THREADS=1 for every rank, yet only rank 0 writes the cloud.  B200D_PLY_THREADS overrides.
    int T = 1;
#ifdef _OPENMP
    T = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("B200D_PLY_THREADS")) T = atoi(e);
    if (T < 1) T = 1;
    if (T > 64) T = 64;
#endif
    const int64_t BLOCK = 1 << 22;                       // vertices formatted per round (bounds the staging memory)
    std::vector<std::vector<char>> bufs((size_t)T);
    std::vector<size_t> used((size_t)T);
    std::vector<off_t> at((size_t)T);
    std::atomic<bool> wbad(false);
    for (int64_t b0 = 0; b0 < n && !bad; b0 += BLOCK) {
        const int64_t bn = (n - b0 < BLOCK) ? n - b0 : BLOCK;
        const int64_t per = (bn + T - 1) / T;
#pragma omp parallel num_threads(T)
        {
#pragma omp for schedule(static, 1)
            for (int t = 0; t < T; ++t) {
                const int64_t i0 = b0 + (int64_t)t * per, i1 = (i0 + per < b0 + bn) ? i0 + per : b0 + bn;
                std::vector<char> &buf = bufs[(size_t)t];
                size_t pos = 0;
                for (int64_t i = i0; i < i1; ++i) {
                    pos = fmt_coord(buf, pos, points[3 * i]);
                    pos = fmt_coord(buf, pos, points[3 * i + 1]);
                    pos = fmt_coord(buf, pos, points[3 * i + 2]);
                    if (buf.size() < pos + 16) buf.resize(buf.size() * 2 + (1 << 16));
                    char *p = buf.data() + pos;
                    p = fmt_u64(p, colors[3 * i]);     *p++ = ' ';
                    p = fmt_u64(p, colors[3 * i + 1]); *p++ = ' ';
                    p = fmt_u64(p, colors[3 * i + 2]); *p++ = '\n';
                    pos = (size_t)(p - buf.data());
                }
                used[(size_t)t] = pos;
            }
            // implicit barrier: all run lengths are known; one thread lays the runs out, then everybody writes its own
#pragma omp single
            for (int t = 0; t < T; ++t) { at[(size_t)t] = filepos; filepos += (off_t)used[(size_t)t]; }
#pragma omp for schedule(static, 1)
            for (int t = 0; t < T; ++t)
                if (used[(size_t)t] && !write_all_at(fd, bufs[(size_t)t].data(), used[(size_t)t], at[(size_t)t])) wbad = true;
        }
        bad = bad || wbad.load();
    }
    if (close(fd) != 0 || bad) return fail(B200D_ERR_ARG, "write failed");
    return B200D_OK;
}

// count-only pass of the fusion mask (same mask as b200d_mask_backproject): lets a caller size one exact output buffer
// for many views with a single host synchronisation
int b200d_mask_count(const float *d_depth, const float *d_conf, int H, int W, float min_conf, int mode, int64_t *d_count,
                     void *d_scratch, void *stream) {
    if (H < 1 || W < 1) return fail(B200D_ERR_ARG, "empty image");
    if (!d_depth || !d_conf || !d_count || !d_scratch) return fail(B200D_ERR_ARG, "null pointer");
    if (mode != 0 && mode != 1) return fail(B200D_ERR_ARG, "mode must be 0 (PatchMatch) or 1 (plane sweep)");
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    a.depth = d_depth; a.conf = d_conf; a.H = H; a.W = W; a.min_conf = min_conf; a.mode = mode; a.count = d_count;
    const size_t chunks = ((size_t)H * W + FUSE_CHUNK - 1) / FUSE_CHUNK;
    a.n_chunks = (int)chunks;
    a.chunk_counts = (uint32_t *)d_scratch;
    a.chunk_offsets = (int64_t *)((char *)d_scratch + align256(chunks * sizeof(uint32_t)));
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(4100, st);
    fuse_count_kernel<<<(unsigned)chunks, FUSE_NT, 0, st>>>(a);
    fuse_scan_kernel<<<1, 1024, 0, st>>>(a);
    return cuda_check(cudaGetLastError(), "fuse count kernels launch");
}

int b200d_mask_backproject(const float *d_depth, const float *d_conf, const uint8_t *d_bgr, int H, int W,
                           const double *Kinv64, const double *R64, const double *t64, float min_conf, int mode,
                           double *d_points, uint8_t *d_colors, int64_t *d_count, void *d_scratch, void *stream) {
    if (H < 1 || W < 1) return fail(B200D_ERR_ARG, "empty image");
    if (!d_depth || !d_conf || !d_bgr || !Kinv64 || !R64 || !t64 || !d_points || !d_colors || !d_count || !d_scratch)
        return fail(B200D_ERR_ARG, "null pointer");
    if (mode != 0 && mode != 1) return fail(B200D_ERR_ARG, "mode must be 0 (PatchMatch) or 1 (plane sweep)");
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
    ProfScope prof(4000, st);
    fuse_count_kernel<<<(unsigned)chunks, FUSE_NT, 0, st>>>(a);
    fuse_scan_kernel<<<1, 1024, 0, st>>>(a);
    fuse_write_kernel<<<(unsigned)chunks, FUSE_NT, 0, st>>>(a);
    return cuda_check(cudaGetLastError(), "fuse kernels launch");
}

}  // extern "C"
