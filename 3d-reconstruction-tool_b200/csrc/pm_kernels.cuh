// pm_kernels.cuh -- fused reproject + bilinear sample + windowed NCC (+ hypothesis select).
//
// One kernel template serves every evaluation of the PatchMatch path
// (reference: src/core/mvs_patchmatch.py:323-534):
//
//   M_STATS   mean1 = box(r), var1 = box(r*r) - mean1^2            (:403,:406, per view)
//   M_COST    cost map of a given depth map                        (_compute_patch_cost :323-390)
//   M_CONF    number of sources with ncc > 0.6 at a depth map      (_compute_confidence :493-534)
//   M_SHIFT   evaluate the shifted-neighbour hypothesis and select (_spatial_propagation :415-457)
//   M_REFINE  evaluate a random perturbation and select            (_random_refinement :459-491)
//
// Work decomposition (per CTA): a TW x TH tile of reference pixels plus a K/2 halo.
//   phase 0  candidate depth for every pixel of tile+halo            -> smem
//   phase A  (per source) reproject + bilinear sample tile+halo      -> smem  s, r*s
//   phase B  (per source) K x K window sums of s, s*s, r*s for the PX x QY outputs each thread
//            owns: register-tiled sliding window, row-major per-tap FMA chain (bit-exact with
//            the oneDNN box filter of the reference), then NCC and aggregation
//   epilogue average / +inf rule, strict `<` selection, state write
//
// The sampled images never leave shared memory: per evaluation HBM sees only the state
// (depth, normal, best) and the L2-resident grays.
#pragma once
#include "device_math.cuh"

namespace b200d {

constexpr int TW = 64, TH = 32;          // tile of output pixels per CTA
constexpr int PX = 4, QY = 2;            // outputs per thread (PX wide, QY tall)
constexpr int NTX = TW / PX, NTY = TH / QY, NT = NTX * NTY;   // 16 x 16 = 256 threads

enum PmMode { M_STATS = 0, M_COST = 1, M_CONF = 2, M_SHIFT = 3, M_REFINE = 4 };

struct PmArgs {
    const float *ref, *mean1, *var1;
    const float *src[B200D_MAX_SRC];
    b200d_cams cams;
    int H, W;
    const float *depth_in, *normal_in, *best_in;
    float *depth_out, *normal_out, *best_out;
    float *out, *out2;                   // cost / confidence / mean1 ; var1
    int dy, dx;                          // M_SHIFT neighbour offset
    float depth_min, depth_max, dr, nr;  // clamp range, refinement ranges of this iteration
    const float *rand_u, *rand_n;        // injected draws of this evaluation (or NULL -> Philox)
    uint64_t seed;
    uint32_t view_id, eval_id;
};

template <int K>
struct Geo {
    static constexpr int HALF = K / 2;
    static constexpr int RW = TW + 2 * HALF;            // region width (valid columns)
    static constexpr int RH = TH + 2 * HALF;            // region height
    static constexpr int RWP = (RW + 3) / 4 * 4;        // padded row stride (float4-aligned rows)
    static constexpr int NL = PX + K - 1;               // floats a thread loads per window row
    static constexpr size_t smem_floats = 4 * (size_t)RH * RWP;   // r, cand depth, s, r*s
    static constexpr size_t smem_bytes = smem_floats * sizeof(float) + (size_t)TH * TW;   // + valid flags
};

// Load NL consecutive floats starting at a 16-byte aligned shared-memory address.
template <int NL>
__device__ __forceinline__ void load_row(const float *__restrict__ p, float (&v)[NL]) {
    static_assert(NL % 2 == 0, "window row length must be even");
#pragma unroll
    for (int i = 0; i + 4 <= NL; i += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
    if (NL % 4) {
        const float2 t = *reinterpret_cast<const float2 *>(p + NL - 2);
        v[NL - 2] = t.x; v[NL - 1] = t.y;
    }
}

// One window row: taps j = 0..K-1 in ascending order for every output (row-major chain).
template <int K, bool DO0, bool DO1>
__device__ __forceinline__ void box_row(const float *__restrict__ srow, const float *__restrict__ rsrow,
                                        float (&as)[QY][PX], float (&ass)[QY][PX], float (&ars)[QY][PX]) {
    constexpr int NL = Geo<K>::NL;
    constexpr float w = 1.0f / (float)(K * K);           // ones(k,k)/(k*k): fp32 RN(1/k^2)
    float sv[NL], rv[NL], s2[NL];
    load_row<NL>(srow, sv);
    load_row<NL>(rsrow, rv);
#pragma unroll
    for (int i = 0; i < NL; ++i) s2[i] = mul(sv[i], sv[i]);
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            if (DO0) {
                as[0][p] = fma_(sv[p + j], w, as[0][p]);
                ass[0][p] = fma_(s2[p + j], w, ass[0][p]);
                ars[0][p] = fma_(rv[p + j], w, ars[0][p]);
            }
            if (DO1) {
                as[1][p] = fma_(sv[p + j], w, as[1][p]);
                ass[1][p] = fma_(s2[p + j], w, ass[1][p]);
                ars[1][p] = fma_(rv[p + j], w, ars[1][p]);
            }
        }
}

// Window sums for the thread's PX x QY outputs.  s_s / s_rs: region arrays, row stride RWP.
template <int K>
__device__ __forceinline__ void box_sums(const float *__restrict__ s_s, const float *__restrict__ s_rs, int tx, int ty,
                                         float (&as)[QY][PX], float (&ass)[QY][PX], float (&ars)[QY][PX]) {
    static_assert(QY == 2, "row peeling below assumes two output rows per thread");
    constexpr int RWP = Geo<K>::RWP;
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) as[q][p] = ass[q][p] = ars[q][p] = 0.0f;
    const int base = (ty * QY) * RWP + tx * PX;
    box_row<K, true, false>(s_s + base, s_rs + base, as, ass, ars);
#pragma unroll 1
    for (int r = 1; r < K; ++r) box_row<K, true, true>(s_s + base + r * RWP, s_rs + base + r * RWP, as, ass, ars);
    box_row<K, false, true>(s_s + base + K * RWP, s_rs + base + K * RWP, as, ass, ars);
}

template <int K, int MODE>
__global__ void __launch_bounds__(NT, 2) pm_eval_kernel(const __grid_constant__ PmArgs a) {
    using G = Geo<K>;
    constexpr int HALF = G::HALF, RW = G::RW, RH = G::RH, RWP = G::RWP;
    extern __shared__ __align__(16) float smem[];
    float *s_r = smem;
    float *s_cd = s_r + RH * RWP;
    float *s_s = s_cd + RH * RWP;
    float *s_rs = s_s + RH * RWP;
    unsigned char *s_valid = reinterpret_cast<unsigned char *>(s_rs + RH * RWP);

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = a.H, W = a.W;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);

    // ---- phase 0: reference tile + candidate depth for tile + halo
    for (int i = tid; i < RH * RW; i += NT) {
        const int ry = i / RW, rx = i - ry * RW;
        const int gx = x0 - HALF + rx, gy = y0 - HALF + ry;
        const bool in = (gx >= 0) && (gx < W) && (gy >= 0) && (gy < H);
        float r = 0.0f, cd = 0.0f;
        if (in) {
            const size_t g = (size_t)gy * W + gx;
            r = __ldg(a.ref + g);
            if (MODE == M_COST || MODE == M_CONF) {
                cd = __ldg(a.depth_in + g);
            } else if (MODE == M_SHIFT) {
                // depth_neighbor = F.pad(shifted slice, value=depth_min)  (mvs_patchmatch.py:429-442)
                const int nx = gx + a.dx, ny = gy + a.dy;
                cd = (nx >= 0 && nx < W && ny >= 0 && ny < H) ? __ldg(a.depth_in + (size_t)ny * W + nx) : a.depth_min;
            } else if (MODE == M_REFINE) {
                float u;
                if (a.rand_u) {
                    u = __ldg(a.rand_u + g);
                } else {
                    const uint4 rn = philox4x32_10(make_uint4((uint32_t)g, a.eval_id, a.view_id, 0u),
                                                   make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                    u = u01(rn.x);
                }
                // depth_delta = (rand*2 - 1) * depth_range ; clamp(depth + delta)   (:471-472)
                const float delta = mul(sub(mul(u, 2.0f), 1.0f), a.dr);
                float dn = add(__ldg(a.depth_in + g), delta);
                dn = dn < a.depth_min ? a.depth_min : dn;
                dn = dn > a.depth_max ? a.depth_max : dn;
                cd = dn;
            }
        }
        s_r[ry * RWP + rx] = r;
        s_cd[ry * RWP + rx] = cd;
        if (MODE == M_STATS) {                     // "sampled image" = the reference itself
            s_s[ry * RWP + rx] = r;
            s_rs[ry * RWP + rx] = mul(r, r);
        }
    }
    if constexpr (RWP > RW) {                      // zero the padding columns once (never rewritten)
        for (int i = tid; i < RH * (RWP - RW); i += NT) {
            const int ry = i / (RWP - RW), rx = RW + i % (RWP - RW);
            s_s[ry * RWP + rx] = 0.0f;
            s_rs[ry * RWP + rx] = 0.0f;
        }
    }
    __syncthreads();

    float as[QY][PX], ass[QY][PX], ars[QY][PX];

    if constexpr (MODE == M_STATS) {
        box_sums<K>(s_s, s_rs, tx, ty, as, ass, ars);
#pragma unroll
        for (int q = 0; q < QY; ++q)
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const int gx = x0 + tx * PX + p, gy = y0 + ty * QY + q;
                if (gx < W && gy < H) {
                    const size_t g = (size_t)gy * W + gx;
                    a.out[g] = as[q][p];
                    a.out2[g] = sub(ars[q][p], mul(as[q][p], as[q][p]));
                }
            }
        return;
    }

    float total[QY][PX], cnt[QY][PX];
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) total[q][p] = cnt[q][p] = 0.0f;

    // validity bounds: strict patch-inside for the cost (:362-363), image-inside for confidence (:516-517)
    const float lo = (MODE == M_CONF) ? 0.0f : (float)HALF;
    const float hix = (MODE == M_CONF) ? (float)W : (float)(W - HALF);
    const float hiy = (MODE == M_CONF) ? (float)H : (float)(H - HALF);

    const int S = a.cams.n_src;
    for (int s = 0; s < S; ++s) {
        const float *__restrict__ img = a.src[s];
        // ---- phase A: reproject + sample tile + halo
        for (int i = tid; i < RH * RW; i += NT) {
            const int ry = i / RW, rx = i - ry * RW;
            const int gx = x0 - HALF + rx, gy = y0 - HALF + ry;
            float sv = 0.0f, rsv = 0.0f;
            if ((gx >= 0) && (gx < W) && (gy >= 0) && (gy < H)) {
                float pw[3], u, v, z;
                backproject_world(a.cams, (float)gx, (float)gy, s_cd[ry * RWP + rx], pw);
                project_src(a.cams, s, pw, u, v, z);
                sv = sample_bilinear(img, W, wm1, hm1, u, v);
                rsv = mul(s_r[ry * RWP + rx], sv);
                const int ox = rx - HALF, oy = ry - HALF;
                if (ox >= 0 && ox < TW && oy >= 0 && oy < TH)
                    s_valid[oy * TW + ox] = (z > 0.1f) && (u >= lo) && (u < hix) && (v >= lo) && (v < hiy);
            }
            s_s[ry * RWP + rx] = sv;
            s_rs[ry * RWP + rx] = rsv;
        }
        __syncthreads();
        // ---- phase B: window sums + NCC for this thread's outputs
        box_sums<K>(s_s, s_rs, tx, ty, as, ass, ars);
#pragma unroll
        for (int q = 0; q < QY; ++q)
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const int ox = tx * PX + p, oy = ty * QY + q;
                const int gx = x0 + ox, gy = y0 + oy;
                if (gx < W && gy < H) {
                    const size_t g = (size_t)gy * W + gx;
                    const float m1 = __ldg(a.mean1 + g), v1 = __ldg(a.var1 + g);
                    const float m2 = as[q][p];
                    const float v2 = sub(ass[q][p], mul(m2, m2));
                    const float cov = sub(ars[q][p], mul(m1, m2));
                    const float ncc = div_(cov, add(__fsqrt_rn(mul(v1, v2)), 1e-8f));     // :410
                    const float cost = sub(1.0f, ncc);
                    const bool valid = s_valid[oy * TW + ox] != 0;
                    if (MODE == M_CONF) {
                        const float ncc2 = sub(1.0f, cost);                               // :530
                        if (valid && (ncc2 > 0.6f)) total[q][p] = add(total[q][p], 1.0f);
                    } else {
                        total[q][p] = add(total[q][p], valid ? cost : 0.0f);              // :383
                        cnt[q][p] = add(cnt[q][p], valid ? 1.0f : 0.0f);
                    }
                }
            }
        __syncthreads();
    }

    // ---- epilogue
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int ox = tx * PX + p, oy = ty * QY + q;
            const int gx = x0 + ox, gy = y0 + oy;
            if (gx >= W || gy >= H) continue;
            const size_t g = (size_t)gy * W + gx;
            if (MODE == M_CONF) {
                a.out[g] = total[q][p];
                continue;
            }
            const float avg = div_(total[q][p], add(cnt[q][p], 1e-8f));                   // :387
            const float cost = (cnt[q][p] >= 2.0f) ? avg : __int_as_float(0x7f800000);   // :388
            if (a.out) a.out[g] = cost;
            if (MODE == M_COST) continue;
            const float best = a.best_in[g];
            const bool better = cost < best;                                              // strict; NaN never wins
            const float cd = s_cd[(oy + HALF) * RWP + ox + HALF];
            float n0, n1, n2;
            if (better) {
                if (MODE == M_SHIFT) {
                    const int nx = gx + a.dx, ny = gy + a.dy;
                    if (nx >= 0 && nx < W && ny >= 0 && ny < H) {
                        const float *np = a.normal_in + 3 * ((size_t)ny * W + nx);
                        n0 = np[0]; n1 = np[1]; n2 = np[2];
                    } else {
                        n0 = n1 = n2 = 0.0f;                                              // F.pad zero fill
                    }
                } else {
                    float r0, r1, r2;
                    if (a.rand_n) {
                        const float *rp = a.rand_n + 3 * g;
                        r0 = __ldg(rp); r1 = __ldg(rp + 1); r2 = __ldg(rp + 2);
                    } else {
                        const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                        const uint4 ra = philox4x32_10(make_uint4((uint32_t)g, a.eval_id, a.view_id, 1u), key);
                        const uint4 rb = philox4x32_10(make_uint4((uint32_t)g, a.eval_id, a.view_id, 2u), key);
                        r0 = gauss(ra.x, ra.y); r1 = gauss(ra.z, ra.w); r2 = gauss(rb.x, rb.y);
                    }
                    // normal_new = F.normalize(normal + randn*normal_range)  (:475-476)
                    const float *np = a.normal_in + 3 * g;
                    const float c0 = add(np[0], mul(r0, a.nr));
                    const float c1 = add(np[1], mul(r1, a.nr));
                    const float c2 = add(np[2], mul(r2, a.nr));
                    float nn = __fsqrt_rn(add(add(mul(c0, c0), mul(c1, c1)), mul(c2, c2)));
                    nn = nn < 1e-12f ? 1e-12f : nn;
                    n0 = div_(c0, nn); n1 = div_(c1, nn); n2 = div_(c2, nn);
                }
            } else if (a.normal_out != a.normal_in) {
                const float *np = a.normal_in + 3 * g;
                n0 = np[0]; n1 = np[1]; n2 = np[2];
            }
            if (better || a.normal_out != a.normal_in) {
                float *no = a.normal_out + 3 * g;
                no[0] = n0; no[1] = n1; no[2] = n2;
            }
            a.depth_out[g] = better ? cd : a.depth_in[g];
            if (better || a.best_out != a.best_in) a.best_out[g] = better ? cost : best;
        }
}

}  // namespace b200d
