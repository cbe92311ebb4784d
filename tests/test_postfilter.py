"""CPU: the device-side post filters (b200dense/postfilter.py) against NumPy run the way the reference runs it
(mvs_patchmatch.py:572-588, dense_stereo.py:475-492).  Same code path as on CUDA tensors; here on CPU tensors."""
import numpy as np
import pytest
import torch

from b200dense.postfilter import _percentile_linear, filter_points, voxel_down_sample


def ref_filter_points(points, colors):
    centroid = np.median(points, axis=0)
    distances = np.linalg.norm(points - centroid, axis=1)
    mask = distances < np.percentile(distances, 95)
    points, colors = points[mask], colors[mask]
    vi = np.floor(points / 0.01).astype(np.int64)
    keys = vi[:, 0] * 1000000000 + vi[:, 1] * 1000000 + vi[:, 2]
    _, idx = np.unique(keys, return_index=True)
    return points[idx], colors[idx]


def ref_voxel(points, colors, voxel):
    vi = np.floor(points / voxel).astype(np.int32)
    keys = vi[:, 0].astype(np.int64) * 1000000000 + vi[:, 1].astype(np.int64) * 1000000 + vi[:, 2].astype(np.int64)
    _, idx = np.unique(keys, return_index=True)
    return points[idx], colors[idx]


@pytest.mark.parametrize('n', [1, 2, 7, 100, 1001, 50000])
def test_percentile_matches_numpy(n):
    rng = np.random.default_rng(n)
    v = np.sort(rng.normal(size=n) ** 2)
    for q in (95, 50, 1, 99.5):
        assert float(_percentile_linear(torch.from_numpy(v), q)) == float(np.percentile(v, q))


@pytest.mark.parametrize('n,seed', [(3, 0), (10, 1), (4999, 2), (60000, 3)])
def test_filter_points_identical_to_numpy(n, seed):
    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(n, 3)) * 0.3
    pts[: n // 50] += 40.0                                       # outliers
    pts = np.concatenate([pts, pts[: n // 3] + 1e-4])            # duplicates inside voxels
    col = rng.integers(0, 255, size=(len(pts), 3)).astype(np.uint8)
    ep, ec = ref_filter_points(pts, col)
    gp, gc = filter_points(torch.from_numpy(pts), torch.from_numpy(col))
    assert np.array_equal(gp.numpy(), ep) and np.array_equal(gc.numpy(), ec)


def test_voxel_down_sample_identical_to_numpy():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-3, 3, size=(30000, 3))
    col = rng.integers(0, 255, size=(30000, 3)).astype(np.uint8)
    ep, ec = ref_voxel(pts, col, 0.02)
    gp, gc = voxel_down_sample(torch.from_numpy(pts), torch.from_numpy(col), 0.02)
    assert np.array_equal(gp.numpy(), ep) and np.array_equal(gc.numpy(), ec)
    e, _ = voxel_down_sample(torch.zeros((0, 3), dtype=torch.float64), torch.zeros((0, 3), dtype=torch.uint8), 0.02)
    assert e.shape == (0, 3)
