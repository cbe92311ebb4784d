"""GPU (-m gpu): the k-NN mean-distance kernel against scikit-learn (the reference's own call, dense_stereo.py:456-460),
bit for bit, and the device outlier filter against the reference's host code."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_filter_outliers(points, colors, k=20, std_ratio=2.0):
    from sklearn.neighbors import NearestNeighbors
    nn = NearestNeighbors(n_neighbors=k).fit(points)
    distances, _ = nn.kneighbors(points)
    mean_dist = np.mean(distances[:, 1:], axis=1)
    mask = mean_dist < np.mean(mean_dist) + std_ratio * np.std(mean_dist)
    return points[mask], colors[mask], mean_dist


@pytest.mark.parametrize('n,k', [(21, 20), (1000, 20), (20011, 20), (5000, 8), (3000, 3)])
def test_knn_mean_dist_bitexact_vs_sklearn(engine, n, k):
    from sklearn.neighbors import NearestNeighbors
    rng = np.random.default_rng(n)
    pts = rng.normal(size=(n, 3)) * 2.0
    pts[: n // 20] = pts[n // 20: 2 * (n // 20)]                 # exact duplicates (zero distances)
    d, _ = NearestNeighbors(n_neighbors=k).fit(pts).kneighbors(pts)
    exp = np.mean(d[:, 1:], axis=1)
    got = engine.knn_mean_dist(engine.to_device(pts, torch.float64), k).cpu().numpy()
    if k >= n // 2:
        # scikit-learn answers tiny clouds (n_neighbors >= n // 2, i.e. n <= 41 at k = 20) by brute force through the
        # |x|^2 - 2xy + |y|^2 expansion, whose distances are 1e-15 off the exact ones its KD-tree (and the kernel) return
        np.testing.assert_allclose(got, exp, rtol=1e-13, atol=1e-14)
    else:
        assert np.array_equal(got, exp)


def test_device_outlier_filter_equals_reference(engine):
    from b200dense.postfilter import filter_outliers
    rng = np.random.default_rng(1)
    pts = np.concatenate([rng.normal(size=(30000, 3)), rng.normal(size=(300, 3)) * 8 + 20])
    col = rng.integers(0, 255, size=(len(pts), 3)).astype(np.uint8)
    ep, ec, _ = ref_filter_outliers(pts, col)
    gp, gc = filter_outliers(engine.to_device(pts, torch.float64), engine.to_device(col, torch.uint8), engine)
    assert np.array_equal(gp.cpu().numpy(), ep) and np.array_equal(gc.cpu().numpy(), ec)
    small = engine.to_device(pts[:10], torch.float64)
    sp, _ = filter_outliers(small, engine.to_device(col[:10], torch.uint8), engine)
    assert sp.shape[0] == 10                                       # fewer than k+1 points: returned unchanged
