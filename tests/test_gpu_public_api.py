"""GPU (-m gpu): the reference-facing classes end to end against `reconstruct()` outputs of the
unmodified reference (tests/golden/reconstruct_wavy_96x72.npz).  With rng='torch' the PatchMatch
class draws its random numbers on the host in the reference's call order, so under the same
torch.manual_seed the whole pipeline (depth maps -> masks -> fusion -> post filters) reproduces the
reference's point cloud: same count, same order, same colours; coordinates to 1e-9."""
import numpy as np
import pytest
import torch

from helpers import load
from core.camera import Camera, CameraPose
from core.dense_stereo import DenseStereoReconstructor, create_combined_dense_cloud
from core.mvs_patchmatch import PatchMatchMVS

pytestmark = pytest.mark.gpu


def _scene(g):
    n = len(g['sc_R'])
    images = [{'image': g['sc_bgr'][i]} for i in range(n)]
    poses = {i: CameraPose(R=g['sc_R'][i], t=g['sc_t'][i]) for i in range(n)}
    return Camera(K=g['sc_K'].copy(), dist=np.zeros(5)), images, poses


@pytest.mark.parametrize('scale', [1.0, 0.5])
def test_plane_sweep_reconstruct_matches_reference(scale):
    g = load('reconstruct_wavy_96x72.npz')
    cam, images, poses = _scene(g)
    st = DenseStereoReconstructor(cam, scale=scale, num_depths=24, patch_size=5)
    p, c = st.reconstruct(images, poses, max_pairs=30)
    ep, ec = g[f'stereo_s{scale}_points'], g[f'stereo_s{scale}_colors']
    assert p.shape == ep.shape
    assert np.array_equal(c, ec)
    np.testing.assert_allclose(p, ep, rtol=0, atol=1e-9)


@pytest.mark.parametrize('scale', [1.0, 0.5])
def test_patchmatch_reconstruct_matches_reference(scale):
    g = load('reconstruct_wavy_96x72.npz')
    cam, images, poses = _scene(g)
    pm = PatchMatchMVS(cam, scale=scale, patch_size=7, num_iterations=2, num_samples=3, min_views=2, rng='torch')
    torch.manual_seed(11)
    p, c = pm.reconstruct(images, poses, sparse_points=g['sparse'])
    assert [pm.depth_min, pm.depth_max] == list(g[f'mvs_s{scale}_range'])
    ep, ec = g[f'mvs_s{scale}_points'], g[f'mvs_s{scale}_colors']
    assert p.shape == ep.shape
    assert np.array_equal(c, ec)
    np.testing.assert_allclose(p, ep, rtol=0, atol=1e-9)


def test_patchmatch_philox_is_deterministic_and_plausible():
    """Production RNG: same seed -> identical maps; different seed -> different maps; on an easy scene
    most confident pixels land near the true depth."""
    from b200dense.scenes import make_scene
    sc = make_scene('plane', 5, 192, 144, seed=8)
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
    out = []
    for seed in (1, 1, 2):
        pm = PatchMatchMVS(cam, scale=1.0, patch_size=7, num_iterations=3, num_samples=8, seed=seed)
        pm.depth_min, pm.depth_max = sc.depth_range
        proc = pm._prepare_images(sc.images, sorted(poses))
        out.append(pm._patchmatch_cuda(2, [1, 3, 0, 4], proc, poses))
    assert np.array_equal(out[0].depth, out[1].depth) and np.array_equal(out[0].normal, out[1].normal)
    assert not np.array_equal(out[0].depth, out[2].depth)
    d = out[0]
    assert np.isfinite(d.depth).all() and d.depth.min() >= sc.depth_range[0] and d.depth.max() <= sc.depth_range[1]
    nrm = np.linalg.norm(d.normal, axis=-1)
    assert np.all((np.abs(nrm - 1) < 1e-5) | (nrm == 0))
    conf = d.confidence >= 3
    assert conf.sum() > 100
    rel = np.abs(d.depth[conf] - sc.depth_gt[2][conf]) / sc.depth_gt[2][conf]
    assert np.median(rel) < 0.05


def test_fewer_than_two_neighbours_skips_view_and_wrapper():
    g = load('reconstruct_wavy_96x72.npz')
    cam, images, poses = _scene(g)
    pm = PatchMatchMVS(cam, scale=0.5, patch_size=5, num_iterations=1, num_samples=1, num_src_views=1)
    p, c = pm.reconstruct(images, poses)                      # every view has < 2 sources -> nothing fused
    assert p.shape == (0, 3) and c.shape == (0, 3)
    p, c = create_combined_dense_cloud(cam, images, poses, use_stereo=False)
    assert len(p) == 0
