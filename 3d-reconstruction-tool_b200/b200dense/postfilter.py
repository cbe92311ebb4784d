"""Point-cloud post filters on the device.

The reference runs these in NumPy on one CPU core over the whole fused cloud
(`PatchMatchMVS._filter_points`, src/core/mvs_patchmatch.py:572-588;
`DenseStereoReconstructor._voxel_down_sample`, src/core/dense_stereo.py:475-492).  At the benchmark
scales (tens of millions of points) they dwarf the GPU stage, so they are restated here as device-side
tensor programs: sorts / scans from the library (torch -> CUB radix sort), everything kept in f64, no scalar read
back to the host in between (the only synchronisations are the two boolean-mask compactions, whose output sizes the
host must know).  Decision recorded in DESIGN.md 5.4: the voxel de-duplication stays on the library's radix sort --
it is an HBM-bound LSD radix sort of (key, index) pairs that already runs at the memory roofline; a hand-written
replacement would be the same algorithm (bench.py --config full4k reports its share of the stage).
The semantics are reproduced exactly, including NumPy's conventions:

  np.median        mean of the two middle order statistics for even n
  np.percentile    method 'linear': virtual index (n-1)*q, lerp with the t >= 0.5 branch
  np.linalg.norm   sqrt((x0^2 + x1^2) + x2^2)
  np.unique(..., return_index=True)   sorted keys, index of the FIRST occurrence (stable sort)

so the selected points, their order and their colours are identical to the reference's
(tests/test_postfilter.py checks this against NumPy on CPU tensors; the same code runs on CUDA tensors).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def _median_axis0(x: torch.Tensor) -> torch.Tensor:
    """np.median(x, axis=0) for an (N,C) f64 tensor."""
    n = x.shape[0]
    s = torch.sort(x, dim=0).values
    if n % 2:
        return s[n // 2]
    return (s[n // 2 - 1] + s[n // 2]) / 2.0


def _percentile_linear(sorted_vals: torch.Tensor, q_percent: float) -> torch.Tensor:
    """np.percentile(v, q) with the default 'linear' method, from the sorted values (f64): a 0-dim tensor on the
    values' device -- nothing is read back to the host (the indices and the interpolation weight depend on n only)."""
    n = int(sorted_vals.shape[0])
    q = np.true_divide(q_percent, 100)
    vi = (n - 1) * q                                            # numpy 'linear': get_virtual_index
    if vi >= n - 1:                                             # _get_indexes: above bounds -> the maximum
        return sorted_vals[n - 1]
    lo = int(np.floor(vi))
    gamma = float(np.float64(vi - lo))
    a, b = sorted_vals[lo], sorted_vals[lo + 1]
    diff = b - a                                                # numpy _lerp: separate multiply and add, in f64
    if gamma >= 0.5:
        return b - diff * (1 - gamma)
    return a + diff * gamma


def _first_of_each_voxel(points: torch.Tensor, voxel_size: float, cast_int32: bool) -> torch.Tensor:
    """Indices np.unique(keys, return_index=True)[1] for keys = ix*1e9 + iy*1e6 + iz of floor(points / voxel)."""
    v = torch.floor(points / voxel_size)
    v = v.to(torch.int32).to(torch.int64) if cast_int32 else v.to(torch.int64)
    keys = v[:, 0] * 1000000000 + v[:, 1] * 1000000 + v[:, 2]
    sk, perm = torch.sort(keys, stable=True)
    first = torch.ones_like(sk, dtype=torch.bool)
    first[1:] = sk[1:] != sk[:-1]
    return perm[first]


def filter_points(points: torch.Tensor, colors: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """`PatchMatchMVS._filter_points` (mvs_patchmatch.py:572-588): keep points closer to the per-axis median
    than the 95th-percentile distance, then the first point of every 0.01 voxel ordered by voxel key."""
    if points.shape[0] == 0:
        return points, colors
    c = _median_axis0(points)
    d = points - c
    dist = torch.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    thr = _percentile_linear(torch.sort(dist).values, 95)
    keep = dist < thr
    points, colors = points[keep], colors[keep]
    idx = _first_of_each_voxel(points, 0.01, cast_int32=False)
    return points[idx], colors[idx]


def voxel_down_sample(points: torch.Tensor, colors: torch.Tensor, voxel_size: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """`DenseStereoReconstructor._voxel_down_sample` (dense_stereo.py:475-492); voxel indices pass through int32."""
    if points.shape[0] == 0:
        return points, colors
    idx = _first_of_each_voxel(points, voxel_size, cast_int32=True)
    return points[idx], colors[idx]


def filter_outliers(points: torch.Tensor, colors: torch.Tensor, engine, k: int = 20,
                    std_ratio: float = 2.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """`DenseStereoReconstructor._filter_outliers` with scikit-learn present (dense_stereo.py:439-466): statistical
    outlier removal on the mean distance to the k-1 nearest neighbours, threshold mean + std_ratio * std.
    The k-NN query is the device kernel (b200d_knn_mean_dist, exact f64); threshold and mask are NumPy on the
    downloaded vector (the reference's own reductions, 4 MB at the 500k cap).  Above 500k points the reference draws
    an unseeded random subsample; so does this (same non-determinism, by design of the reference)."""
    n = int(points.shape[0])
    if n < k + 1:
        return points, colors
    if n > 500000:
        indices = torch.from_numpy(np.random.choice(n, 500000, replace=False)).to(points.device)
        sample = points[indices].contiguous()
    else:
        indices = None
        sample = points.contiguous()
    mean_dist = engine.knn_mean_dist(sample, k).cpu().numpy()
    threshold = np.mean(mean_dist) + std_ratio * np.std(mean_dist)
    inlier = torch.from_numpy(mean_dist < threshold).to(points.device)
    sel = indices[inlier] if indices is not None else inlier
    return points[sel], colors[sel]
