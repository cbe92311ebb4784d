// ncc_kernels.cuh -- windowed NCC over the sampled images + aggregation + hypothesis selection.
//
// Second half of a cost evaluation (reference: `_ncc_cost` / `_compute_ncc_torch` and the callers'
// aggregation and selection: src/core/mvs_patchmatch.py:380-388, 392-413, 452-455, 486-489, 530-532;
// src/core/dense_stereo.py:300-310, 318-347).
//
// Per CTA: a TW x TH tile of reference pixels.  One elected thread streams the (tile + K/2 halo)
// window of the reference image and of each sampled image into shared memory with TMA
// (cp.async.bulk.tensor.2d, double-buffered over the sources, mbarrier completion); out-of-image
// elements are zero-filled by the TMA unit, which is exactly the zero padding of
// F.conv2d(..., padding=K//2).  All threads then run the K x K window sums of s, s*s and r*s for
// their PX x QY outputs as a register-tiled sliding window.  The sums are the reference's
// row-major per-tap FMA chains (bit-exact); two output rows share every tap value,
// so each tap is one packed FFMA2 (fma.rn.f32x2: broadcast tap, immediate 1/K^2, accumulator pair)
// -- half the issue slots of scalar FFMAs, which is what lets the FP32 pipe, not the issue port,
// bound this kernel.
#pragma once
#include <cuda.h>

#include "device_math.cuh"

namespace b200d {

#ifndef B200D_TW
#define B200D_TW 64
#endif
#ifndef B200D_TH
#define B200D_TH 32
#endif
#ifndef B200D_MIN_BLOCKS
#define B200D_MIN_BLOCKS 3
#endif
constexpr int TW = B200D_TW, TH = B200D_TH;                    // output tile per CTA
constexpr int PX = 4, QY = 2;                                  // outputs per thread
constexpr int NTX = TW / PX, NTY = TH / QY, NT = NTX * NTY;    // 16 x 16 threads

enum NccMode { N_STATS = 0, N_COST = 1, N_CONF = 2, N_SHIFT = 3, N_REFINE = 4, N_SWEEP = 5,
               N_SWEEP_DUMP = 6,            // the sweep that also writes the (D,N,H,W) NCC volume (tests only)
               N_VPROP = 7 };               // view propagation (extension): candidate depth + normal maps, adoption mask

struct NccArgs {
    CUtensorMap tm_r;                        // reference gray, dims (W+PAD,H), pixel x at column x+PAD
    CUtensorMap tm_s[B200D_MAX_SRC];         // sampled images, same layout
    int n_src, H, W;
    const float *mean1, *var1;               // reference window statistics
    const uint8_t *valid;                    // per-pixel validity bits from the sampling kernel
    const float *cand;                       // candidate depth evaluated here
    const float *depth_in, *normal_in, *best_in;
    float *depth_out, *normal_out, *best_out;
    float *out, *out2;                       // cost / confidence / mean1 ; var1
    int dy, dx;                              // N_SHIFT neighbour offset (candidate normal)
    int color;                               // N_SHIFT: -1 = every pixel may adopt the candidate (reference schedule);
                                             // 0 / 1 = only pixels with (x + y) & 1 == color (red-black extension)
    const uint8_t *allow;                    // N_VPROP: 1 = this pixel's candidate is a real proposal (may be adopted)
    float nr;                                // N_REFINE normal range
    const float *rand_n;                     // injected N(0,1) (H,W,3) of this evaluation or NULL
    uint64_t seed;
    uint32_t view_id, eval_id;
    // N_SWEEP
    float thresh;
    float t_lo, t_hi;                        // ((1 -+ 1e-5) thresh)^2 ; 0 = always take the exact comparison
    int plane, n_planes;
    const float *depths;
    float *conf, *depth_map, *ncc;
    int32_t *idx;
    // raw views of the TMA sources (bounded plane sweep: exact redo of single outputs straight from L2)
    const float *smap, *rpad;                // sampled maps / reference copy: pixel x at column x + pad, row pitch Wp
    size_t map_stride;
    int Wp, pad;
};

template <int K>
struct Geo {
    static constexpr int HALF = K / 2;
    static constexpr int RW = TW + 2 * HALF;                 // window columns actually used
    static constexpr int RH = TH + 2 * HALF;
    static constexpr int RWP = (RW + 3) / 4 * 4;             // TMA box width / smem row stride
    // TMA (tiled, fp32, no swizzle) faults with an illegal-instruction error unless the box starts on a
    // 16-byte boundary of the row (measured on B200: profiles/microbench/tma_test.cu).  The window starts at
    // x0 - HALF, so every TMA source stores pixel x at column x + PAD with PAD = HALF mod 4 and keeps its
    // first PAD columns zero (they stand for x < 0); out-of-image elements elsewhere are zero-filled by TMA.
    static constexpr int PAD = HALF % 4;
    static constexpr int NL = PX + K - 1;                    // floats per thread per window row
    static constexpr uint32_t tile_bytes = RWP * RH * sizeof(float);
    static constexpr uint32_t buf_bytes = (tile_bytes + 127) / 128 * 128;
    // Keep this small: 3 CTAs x 38 KB selects the 132 KB shared-memory carve-out and leaves ~96 KB of L1 to the
    // sampling kernels of other views that share the SM (a 57 KB variant with per-tile statistics in shared memory
    // forced the 196 KB carve-out and cost 9 % of the concurrent throughput: profiles/r01_notes.md).
    static constexpr size_t smem_bytes = 3 * (size_t)buf_bytes + 64;         // r + 2 stages + barriers
};

// ---- mbarrier / TMA primitives (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, uint64_t *bar, int cx, int cy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(cx), "r"(cy)
        : "memory");
}

// ---- packed fp32x2 arithmetic (each half is one IEEE-RN operation)
__device__ __forceinline__ void fma2_bcast(float &lo, float &hi, float x, float w) {      // {lo,hi} += x*w, x*w
    unsigned long long acc, xx, ww;
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "f"(lo), "f"(hi));
    asm("mov.b64 %0, {%1,%1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(ww));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
}
__device__ __forceinline__ void mul2(float &olo, float &ohi, float alo, float ahi, float blo, float bhi) {
    unsigned long long o, a, b;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(alo), "f"(ahi));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(blo), "f"(bhi));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o) : "l"(a), "l"(b));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(olo), "=f"(ohi) : "l"(o));
}

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

// One window row.  Taps j = 0..K-1 ascending for every output: together with the ascending row
// order in box_sums this is the row-major chain acc = fma(x, 1/K^2, acc) of the reference's conv2d.
template <int K, bool DO0, bool DO1>
__device__ __forceinline__ void box_row(const float *__restrict__ srow, const float *__restrict__ rrow,
                                        float (&as)[QY][PX], float (&ass)[QY][PX], float (&ars)[QY][PX]) {
    constexpr int NL = Geo<K>::NL;
    constexpr float w = 1.0f / (float)(K * K);           // ones(k,k)/(k*k): fp32 RN(1/k^2)
    float sv[NL], rv[NL], s2[NL], rs[NL];
    load_row<NL>(srow, sv);
    load_row<NL>(rrow, rv);
#pragma unroll
    for (int i = 0; i < NL; i += 2) {
        mul2(s2[i], s2[i + 1], sv[i], sv[i + 1], sv[i], sv[i + 1]);      // img2 * img2
        mul2(rs[i], rs[i + 1], rv[i], rv[i + 1], sv[i], sv[i + 1]);      // img1 * img2
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            if (DO0 && DO1) {
                fma2_bcast(as[0][p], as[1][p], sv[p + j], w);
                fma2_bcast(ass[0][p], ass[1][p], s2[p + j], w);
                fma2_bcast(ars[0][p], ars[1][p], rs[p + j], w);
            } else {
                constexpr int q = DO0 ? 0 : 1;
                as[q][p] = fma_(sv[p + j], w, as[q][p]);
                ass[q][p] = fma_(s2[p + j], w, ass[q][p]);
                ars[q][p] = fma_(rs[p + j], w, ars[q][p]);
            }
        }
}

template <int K>
__device__ __forceinline__ void box_sums(const float *__restrict__ s_s, const float *__restrict__ s_r, int tx, int ty,
                                         float (&as)[QY][PX], float (&ass)[QY][PX], float (&ars)[QY][PX]) {
    static_assert(QY == 2, "row peeling below assumes two output rows per thread");
    constexpr int RWP = Geo<K>::RWP;
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) as[q][p] = ass[q][p] = ars[q][p] = 0.0f;
    const int base = (ty * QY) * RWP + tx * PX;
    box_row<K, true, false>(s_s + base, s_r + base, as, ass, ars);
#pragma unroll 2
    for (int r = 1; r < K; ++r) box_row<K, true, true>(s_s + base + r * RWP, s_r + base + r * RWP, as, ass, ars);
    box_row<K, false, true>(s_s + base + K * RWP, s_r + base + K * RWP, as, ass, ars);
}

// ---- N_REFINE: normal_new = F.normalize(normal + randn * normal_range) for the winners (mvs_patchmatch.py:475-476, :486-489).
// In the later evaluations a few per cent of the pixels win, yet nearly every (warp, output slot) has a winner in
// some lane, so running this per output slot would cost the full Philox + Box-Muller + normalise sequence eight
// times per warp.  Instead the warp enumerates its winners (ballot per slot, prefix of the counts) and lane l
// takes winners l, l+32, ...: each winner is processed once, by whichever lane, straight from / to global memory.
// `wins`: bit q*PX+p = output (q,p) of this thread won.  Called by every lane of the warp (exact and separable kernels).
__device__ __forceinline__ void refine_normals_compacted(const NccArgs &a, unsigned wins, int x0, int y0, int tid, int W) {
    const int lane = tid & 31;
    __syncwarp();
    unsigned m_sel = 0, b_sel = 0, total_w = 0;        // mask / first winner index of the slot this lane's winner is in
    // lane-uniform tables would need dynamic indexing; instead each lane keeps what it needs for w = lane + 32*it
    unsigned masks[QY * PX], bases[QY * PX + 1];
    bases[0] = 0;
#pragma unroll
    for (int sl = 0; sl < QY * PX; ++sl) {
        masks[sl] = __ballot_sync(0xffffffffu, (wins >> sl) & 1u);
        bases[sl + 1] = bases[sl] + __popc(masks[sl]);
    }
    total_w = bases[QY * PX];
    const int warp_tid0 = tid & ~31;
    for (unsigned w = lane; w < total_w; w += 32) {
        int sl = 0;
        m_sel = masks[0];
        b_sel = 0;
#pragma unroll
        for (int k = 1; k < QY * PX; ++k)
            if (w >= bases[k]) { sl = k; m_sel = masks[k]; b_sel = bases[k]; }
        const int owner = warp_tid0 + (int)__fns(m_sel, 0, (int)(w - b_sel) + 1);
        const int gx = x0 + (owner % NTX) * PX + sl % PX, gy = y0 + (owner / NTX) * QY + sl / PX;
        const size_t g = (size_t)gy * W + gx;
        float r0, r1, r2;
        if (a.rand_n) {
            const float *rp = a.rand_n + 3 * g;
            r0 = __ldg(rp); r1 = __ldg(rp + 1); r2 = __ldg(rp + 2);
        } else {
            const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            const uint4 ra = philox4x32_10(make_uint4((uint32_t)g, a.eval_id, a.view_id, 1u), key);
            gauss2(ra.x, ra.y, r0, r1);
            r2 = gauss(ra.z, ra.w);
        }
        const float *np = a.normal_in + 3 * g;
        const float c0 = add(np[0], mul(r0, a.nr));
        const float c1 = add(np[1], mul(r1, a.nr));
        const float c2 = add(np[2], mul(r2, a.nr));
        float nn = __fsqrt_rn(add(add(mul(c0, c0), mul(c1, c1)), mul(c2, c2)));
        nn = nn < 1e-12f ? 1e-12f : nn;
        float *no = a.normal_out + 3 * g;
        no[0] = div_(c0, nn); no[1] = div_(c1, nn); no[2] = div_(c2, nn);
    }
}

// Plane sweep: `ncc > thresh` for ncc = cov / sqrt(x), x = var1*var2 + 1e-8 (dense_stereo.py:303, 344-345), decided
// without the square root and the division whenever cov^2 is clearly off thresh^2 * x (t_lo = ((1 - 1e-5) thresh)^2,
// t_hi = ((1 + 1e-5) thresh)^2: the computed quotient is within (1 +- 4u) of cov / sqrt(x) and the squared comparison
// adds 2u, so the 1e-5 margin is ~30x what the implication needs); only the sliver around the
// threshold, zero / negative / non-finite x, and a disabled quick test (t_lo == 0) take the exact path.
__device__ __forceinline__ bool sweep_over(float cov, float x, float thresh, float t_lo, float t_hi) {
    const float c2 = mul(cov, cov);
    const bool cpos = cov > 0.0f, xpos = x > 0.0f;                                  // both false for NaN
    const bool sure_f = !cpos | (xpos & (c2 < mul(x, t_lo)));                        // ncc <= 0, -inf, NaN, or clearly below
    const bool sure_t = cpos & xpos & (c2 > mul(x, t_hi));                           // clearly above
    bool over = sure_t;
    if (!(sure_f | sure_t) | !(t_lo > 0.0f)) over = div_(cov, __fsqrt_rn(x)) > thresh;
    return over;
}

template <int K, int MODE>
__global__ void __launch_bounds__(NT, B200D_MIN_BLOCKS) ncc_kernel(const __grid_constant__ NccArgs a) {
    using G = Geo<K>;
    constexpr int HALF = G::HALF;
    extern __shared__ __align__(128) unsigned char sm[];     // dynamic smem window starts 128-byte aligned
    constexpr int BUF = G::buf_bytes / sizeof(float);
    float *s_r = reinterpret_cast<float *>(sm);              // [0,BUF): reference; [BUF,3*BUF): two sample stages
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 3 * G::buf_bytes);   // [0],[1]: stages ; [2]: reference ; [3],[4]: stage released

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = a.H, W = a.W, S = a.n_src;
    const int cx = x0 - HALF + G::PAD, cy = y0 - HALF;     // cx is a multiple of 4

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        mbar_init(&bar[3], NT / 32);                       // one arrival per warp: "stage 0 may be refilled"
        mbar_init(&bar[4], NT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        if (MODE != N_STATS) {
            mbar_expect_tx(&bar[2], G::tile_bytes);
            tma_load_2d(s_r, &a.tm_r, &bar[2], cx, cy);
        }
        mbar_expect_tx(&bar[0], G::tile_bytes);
        tma_load_2d(s_r + BUF, &a.tm_s[0], &bar[0], cx, cy);
        if (S > 1) {
            mbar_expect_tx(&bar[1], G::tile_bytes);
            tma_load_2d(s_r + 2 * BUF, &a.tm_s[1], &bar[1], cx, cy);
        }
    }

    // per-output constants, issued before the first wait so they are in flight during the TMA
    float m1[QY][PX], v1[QY][PX];
    unsigned vb[QY];                                       // validity bytes of the 4 outputs of a row, packed
#pragma unroll
    for (int q = 0; q < QY; ++q) {
        vb[q] = 0;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int gx = x0 + tx * PX + p, gy = y0 + ty * QY + q;
            const bool in = gx < W && gy < H;
            const size_t g = (size_t)gy * W + gx;
            m1[q][p] = (in && MODE != N_STATS) ? __ldg(a.mean1 + g) : 0.0f;
            v1[q][p] = (in && MODE != N_STATS) ? __ldg(a.var1 + g) : 0.0f;
            vb[q] |= ((in && MODE != N_STATS) ? (unsigned)__ldg(a.valid + g) : 0u) << (8 * p);
        }
    }

    float as[QY][PX], ass[QY][PX], ars[QY][PX], total[QY][PX];
    unsigned cnt = 0;                                      // 4-bit valid-source counters of the 8 outputs
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) total[q][p] = 0.0f;

    const int lane = tid & 31;
    const bool warp_live = y0 + (tid / 32) * (32 / NTX) * QY < H;
    if (MODE != N_STATS) mbar_wait(&bar[2], 0);
    for (int s = 0; s < S; ++s) {
        const int st = s & 1;
        mbar_wait(&bar[st], (s >> 1) & 1);
        // a warp owns QY*2 consecutive tile rows; in the bottom row of tiles some warps lie wholly below the image:
        // they keep taking part in the stage hand-over but have nothing to sum or to select.  The same holds for a
        // (warp, source) none of whose 256 outputs has a valid projection into this source (image borders, hypotheses
        // far off the surface): every use of the NCC below is gated by `valid` (mvs_patchmatch.py:383-384, 530-532;
        // dense_stereo.py:303), so the chains would be computed and thrown away -- skip them (bit-identical results)
        bool src_live = warp_live;
        if (MODE != N_STATS && MODE != N_SWEEP_DUMP)
            src_live = __any_sync(0xffffffffu, (((vb[0] | vb[1]) >> s) & 0x01010101u) != 0u) && warp_live;
        if (src_live) box_sums<K>(s_r + (1 + st) * BUF, (MODE == N_STATS) ? s_r + BUF : s_r, tx, ty, as, ass, ars);
        if (MODE == N_STATS) break;
        if (s + 2 < S) {
            // release the stage: one arrival per warp; the producer thread refills it once all 8 arrived
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[3 + st])) : "memory");
            if (tid == 0) {
                mbar_wait(&bar[3 + st], (s >> 1) & 1);
                mbar_expect_tx(&bar[st], G::tile_bytes);
                tma_load_2d(s_r + (1 + st) * BUF, &a.tm_s[s + 2], &bar[st], cx, cy);
            }
        }
        // PatchMatch modes: ncc = cov / (sqrt(var1*var2) + 1e-8) (mvs_patchmatch.py:410) for the eight outputs, first
        // all of them through the branch-free fast paths (eight independent MUFU + FMA chains the compiler interleaves),
        // then the rare untrusted ones (negative / non-finite / near-denormal operands) again with the IEEE intrinsics
        float nccv[QY][PX];
        if (src_live && MODE != N_STATS && MODE != N_SWEEP && MODE != N_SWEEP_DUMP) {
            unsigned bad = 0;
#pragma unroll
            for (int q = 0; q < QY; ++q)
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const float m2 = as[q][p];
                    const float pv = mul(v1[q][p], sub(ass[q][p], mul(m2, m2)));
                    const float cov = sub(ars[q][p], mul(m1[q][p], m2));
                    bool t_s, t_d;
                    const float den = add(sqrt_fast(pv, t_s), 1e-8f);                   // in [1e-8, ~1] whenever t_s holds
                    nccv[q][p] = div_fast(cov, den, t_d);
                    bad |= (t_s & t_d) ? 0u : 1u << (q * PX + p);
                }
            if (bad) {
#pragma unroll
                for (int q = 0; q < QY; ++q)
#pragma unroll
                    for (int p = 0; p < PX; ++p)
                        if ((bad >> (q * PX + p)) & 1u) {
                            const float m2 = as[q][p];
                            const float pv = mul(v1[q][p], sub(ass[q][p], mul(m2, m2)));
                            nccv[q][p] = div_(sub(ars[q][p], mul(m1[q][p], m2)), add(__fsqrt_rn(pv), 1e-8f));
                        }
            }
        }
        if (src_live)
#pragma unroll
        for (int q = 0; q < QY; ++q)
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const float m1v = m1[q][p], v1v = v1[q][p];
                const float m2 = as[q][p];
                const float v2 = sub(ass[q][p], mul(m2, m2));
                const float cov = sub(ars[q][p], mul(m1v, m2));
                const bool valid = (vb[q] >> (8 * p + s)) & 1u;
                if (MODE == N_SWEEP || MODE == N_SWEEP_DUMP) {
                    // ncc = cov / sqrt(var1*var2 + 1e-8) (dense_stereo.py:344-345) is only ever compared with the
                    // threshold (:303).  The comparison is decided without the square root and the division whenever
                    // cov^2 is clearly off thresh^2 * x (1e-5 margins, ~30x the rounding errors involved: sweep_over);
                    // only the sliver around the threshold (and zero / negative / non-finite x) takes the exact path.
                    const float x = add(mul(v1v, v2), 1e-8f);
                    bool over;
                    if (MODE == N_SWEEP_DUMP) {
                        const float ncc = div_(cov, __fsqrt_rn(x));
                        const int gx = x0 + tx * PX + p, gy = y0 + ty * QY + q;
                        if (gx < W && gy < H) a.ncc[(((size_t)a.plane * S + s) * H + gy) * W + gx] = ncc;
                        over = ncc > a.thresh;
                    } else {
                        over = sweep_over(cov, x, a.thresh, a.t_lo, a.t_hi);
                    }
                    if (over && valid) total[q][p] = add(total[q][p], 1.0f);                  // :303-304
                } else {
                    const float cost = sub(1.0f, nccv[q][p]);
                    if (MODE == N_CONF) {
                        const float ncc2 = sub(1.0f, cost);                                   // :530
                        if (valid && (ncc2 > 0.6f)) total[q][p] = add(total[q][p], 1.0f);
                    } else {
                        total[q][p] = add(total[q][p], valid ? cost : 0.0f);                  // :383
                        cnt += valid ? (1u << (4 * (q * PX + p))) : 0u;
                    }
                }
            }
    }

    // ---- epilogue
    // avg = total / (count + 1e-8) (:387) for the eight outputs, batched through the branch-free division like the NCC
    // above (the divisor is 1e-8, 1, 2, ..., S: always a trusted one)
    float avgv[QY][PX];
    if (MODE == N_COST || MODE == N_SHIFT || MODE == N_REFINE || MODE == N_VPROP) {
        unsigned bad = 0;
#pragma unroll
        for (int q = 0; q < QY; ++q)
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const float cntf = (float)((cnt >> (4 * (q * PX + p))) & 15u);
                bool t;
                avgv[q][p] = div_fast(total[q][p], add(cntf, 1e-8f), t);
                bad |= t ? 0u : 1u << (q * PX + p);
            }
        if (bad) {
#pragma unroll
            for (int q = 0; q < QY; ++q)
#pragma unroll
                for (int p = 0; p < PX; ++p)
                    if ((bad >> (q * PX + p)) & 1u)
                        avgv[q][p] = div_(total[q][p], add((float)((cnt >> (4 * (q * PX + p))) & 15u), 1e-8f));
        }
    }
    unsigned wins = 0;                                     // N_REFINE: bit q*PX+p = output (q,p) of this thread won
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int gx = x0 + tx * PX + p, gy = y0 + ty * QY + q;
            if (gx >= W || gy >= H) continue;
            const size_t g = (size_t)gy * W + gx;
            if (MODE == N_STATS) {
                a.out[g] = as[q][p];
                a.out2[g] = sub(ars[q][p], mul(as[q][p], as[q][p]));
                continue;
            }
            if (MODE == N_CONF) {
                a.out[g] = total[q][p];
                continue;
            }
            if (MODE == N_SWEEP || MODE == N_SWEEP_DUMP) {
                // running first-max over planes (torch.max(dim=0), dense_stereo.py:307)
                const float prev = (a.plane == 0) ? -1.0f : a.conf[g];
                int bi = (a.plane == 0) ? 0 : a.idx[g];
                float bc = prev;
                if (total[q][p] > prev) { bc = total[q][p]; bi = a.plane; }
                if (bc != prev) { a.conf[g] = bc; a.idx[g] = bi; }
                if (a.plane == a.n_planes - 1) a.depth_map[g] = __ldg(a.depths + bi);          // :310
                continue;
            }
            const float cntf = (float)((cnt >> (4 * (q * PX + p))) & 15u);
            const float avg = avgv[q][p];                                                     // :387
            const float cost = (cntf >= 2.0f) ? avg : __int_as_float(0x7f800000);            // :388
            if (a.out) a.out[g] = cost;
            if (MODE == N_COST) continue;
            const float best = a.best_in[g];
            bool better = cost < best;                                                        // strict; NaN never wins
            if (MODE == N_SHIFT && a.color >= 0) better = better && ((gx + gy) & 1) == a.color;
            if (MODE == N_VPROP) better = better && (__ldg(a.allow + g) != 0);
            if (better) a.depth_out[g] = __ldg(a.cand + g);
            else if (a.depth_out != a.depth_in) a.depth_out[g] = a.depth_in[g];
            if (better || a.best_out != a.best_in) a.best_out[g] = better ? cost : best;
            if (MODE == N_REFINE) {
                // the winners' normals are perturbed below, compacted over the warp; a loser only carries its normal
                // over when the evaluation is not in place
                wins |= better ? 1u << (q * PX + p) : 0u;
                if (!better && a.normal_out != a.normal_in) {
                    const float *np = a.normal_in + 3 * g;
                    float *no = a.normal_out + 3 * g;
                    no[0] = np[0]; no[1] = np[1]; no[2] = np[2];
                }
                continue;
            }
            if (MODE == N_VPROP) {
                // the candidate normal map (normal_in) is adopted by the winners; everybody else keeps the state
                if (better) {
                    const float *np = a.normal_in + 3 * g;
                    float *no = a.normal_out + 3 * g;
                    no[0] = np[0]; no[1] = np[1]; no[2] = np[2];
                }
                continue;
            }
            float n0, n1, n2;
            if (better) {
                const int nx = gx + a.dx, ny = gy + a.dy;
                if (nx >= 0 && nx < W && ny >= 0 && ny < H) {
                    const float *np = a.normal_in + 3 * ((size_t)ny * W + nx);
                    n0 = np[0]; n1 = np[1]; n2 = np[2];
                } else {
                    n0 = n1 = n2 = 0.0f;                                                      // F.pad zero fill
                }
            } else if (a.normal_out != a.normal_in) {
                const float *np = a.normal_in + 3 * g;
                n0 = np[0]; n1 = np[1]; n2 = np[2];
            }
            if (better || a.normal_out != a.normal_in) {
                float *no = a.normal_out + 3 * g;
                no[0] = n0; no[1] = n1; no[2] = n2;
            }
        }

    if (MODE == N_REFINE) refine_normals_compacted(a, wins, x0, y0, tid, W);
}

}  // namespace b200d
