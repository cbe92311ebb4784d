// knn_kernels.cuh -- mean distance to the k nearest neighbours, exact f64.
//
// Reference: DenseStereoReconstructor._filter_outliers (src/core/dense_stereo.py:439-473):
//   NearestNeighbors(n_neighbors=20).fit(P).kneighbors(P)  ->  np.mean(distances[:, 1:], axis=1).
// scikit-learn answers this with a KD-tree whose point distances are sqrt(((dx*dx + dy*dy) + dz*dz)) in f64 (no
// FMA), neighbours sorted ascending, the point itself first.  A brute-force scan that keeps the k smallest reduced
// distances per query reproduces those numbers bit for bit (ties only duplicate values), and the mean follows
// NumPy's pairwise order for 19 terms.  N^2 work, FP64-pipe bound: 2.5e11 pairs at the reference's 500k-point cap.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200d {

constexpr int KNN_NT = 128, KNN_Q = 2, KNN_TILE = 256, KNN_KMAX = 32;

__global__ void __launch_bounds__(KNN_NT) knn_mean_dist_kernel(const double *__restrict__ pts, int64_t n, int k,
                                                               double *__restrict__ out) {
    __shared__ double s_p[KNN_TILE * 3];
    double best[KNN_Q][KNN_KMAX];
    double qx[KNN_Q], qy[KNN_Q], qz[KNN_Q], worst[KNN_Q];
    int64_t qi[KNN_Q];
#pragma unroll
    for (int u = 0; u < KNN_Q; ++u) {
        qi[u] = ((int64_t)blockIdx.x * KNN_NT + threadIdx.x) * KNN_Q + u;
        const int64_t j = qi[u] < n ? qi[u] : n - 1;
        qx[u] = pts[3 * j]; qy[u] = pts[3 * j + 1]; qz[u] = pts[3 * j + 2];
        for (int i = 0; i < k; ++i) best[u][i] = __longlong_as_double(0x7ff0000000000000LL);
        worst[u] = best[u][0];
    }
    for (int64_t t0 = 0; t0 < n; t0 += KNN_TILE) {
        const int cntp = (int)((n - t0 < KNN_TILE) ? (n - t0) : KNN_TILE);
        __syncthreads();
        for (int i = threadIdx.x; i < cntp * 3; i += KNN_NT) s_p[i] = pts[3 * t0 + i];
        __syncthreads();
        for (int j = 0; j < cntp; ++j) {
            const double px = s_p[3 * j], py = s_p[3 * j + 1], pz = s_p[3 * j + 2];
#pragma unroll
            for (int u = 0; u < KNN_Q; ++u) {
                const double dx = __dsub_rn(qx[u], px), dy = __dsub_rn(qy[u], py), dz = __dsub_rn(qz[u], pz);
                const double r = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (r < worst[u]) {                       // rare after the first tiles: sorted insertion
                    int i = k - 1;
                    while (i > 0 && best[u][i - 1] > r) {
                        best[u][i] = best[u][i - 1];
                        --i;
                    }
                    best[u][i] = r;
                    worst[u] = best[u][k - 1];
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < KNN_Q; ++u) {
        if (qi[u] >= n) continue;
        // np.mean(distances[1:k]) with NumPy's pairwise order: 8 running sums over blocks of 8, tree, then the tail
        const int m = k - 1;
        double sum;
        if (m < 8) {
            sum = 0.0;
            for (int i = 0; i < m; ++i) sum = __dadd_rn(sum, __dsqrt_rn(best[u][1 + i]));
            if (m > 0) {                                   // NumPy starts from the first element, not from 0.0 (same value)
            }
        } else {
            double r8[8];
            for (int j = 0; j < 8; ++j) r8[j] = __dsqrt_rn(best[u][1 + j]);
            int i = 8;
            for (; i + 8 <= m; i += 8)
                for (int j = 0; j < 8; ++j) r8[j] = __dadd_rn(r8[j], __dsqrt_rn(best[u][1 + i + j]));
            sum = __dadd_rn(__dadd_rn(__dadd_rn(r8[0], r8[1]), __dadd_rn(r8[2], r8[3])),
                            __dadd_rn(__dadd_rn(r8[4], r8[5]), __dadd_rn(r8[6], r8[7])));
            for (; i < m; ++i) sum = __dadd_rn(sum, __dsqrt_rn(best[u][1 + i]));
        }
        out[qi[u]] = __ddiv_rn(sum, (double)m);
    }
}

}  // namespace b200d
