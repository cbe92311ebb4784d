"""GPU (-m gpu): the geometric-consistency EXTENSION (no dackey-wav oracle -- the reference's check is photometric,
SURVEY F6).  The kernel must reproduce, bit for bit, the definition restated in oracle/dense_oracle.c
(orc_geo_consistency: same fp32 operation sequence as the pinned sampling path, K with skew, fp32 Kinv); it is also
cross-checked against an independent PyTorch formulation (borderline pixels may differ there: different op order) and
against ground truth: exact depth maps are consistent in every source that sees the point, corrupted depths are
rejected."""
import numpy as np
import pytest
import torch

from b200dense.engine import kinv32, make_cams
from b200dense.scenes import make_scene

pytestmark = pytest.mark.gpu


def torch_geo(depth_ref, depth_srcs, K, R_ref, t_ref, Rs, ts, px, rel):
    H, W = depth_ref.shape
    dev = depth_ref.device
    f = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)
    K, R_ref, t_ref = f(K), f(R_ref), f(t_ref)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32),
                            torch.arange(W, device=dev, dtype=torch.float32), indexing='ij')
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    d = depth_ref
    pc = torch.stack([(xs - cx) / fx * d, (ys - cy) / fy * d, d], -1) - t_ref
    Xw = pc @ R_ref
    count = torch.zeros_like(d)
    for R, t, ds_map in zip(Rs, ts, depth_srcs):
        R, t = f(R), f(t)
        q = Xw @ R.T + t
        ok = (q[..., 2] > 1e-6) & (d > 0)
        z = torch.where(ok, q[..., 2], torch.ones_like(d))
        us = torch.round(fx * q[..., 0] / z + cx)
        vs = torch.round(fy * q[..., 1] / z + cy)
        ok &= (us >= 0) & (us < W) & (vs >= 0) & (vs < H)
        ui, vi = us.clamp(0, W - 1).long(), vs.clamp(0, H - 1).long()
        ds = ds_map[vi, ui]
        ok &= ds > 0
        qs = torch.stack([(us - cx) / fx * ds, (vs - cy) / fy * ds, ds], -1) - t
        pr = (qs @ R) @ R_ref.T + t_ref
        ok &= pr[..., 2] > 1e-6
        zz = torch.where(pr[..., 2] > 1e-6, pr[..., 2], torch.ones_like(d))
        du = fx * pr[..., 0] / zz + cx - xs
        dv = fy * pr[..., 1] / zz + cy - ys
        ok &= (du * du + dv * dv < px * px) & ((pr[..., 2] - d).abs() < rel * d)
        count += ok.float()
    return count


def test_geo_consistency_matches_definition_and_ground_truth(engine):
    sc = make_scene('wavy', 5, 200, 150, seed=12)
    ref, src = 2, [1, 3, 0, 4]
    R = [sc.poses[i][0] for i in range(5)]
    t = [sc.poses[i][1] for i in range(5)]
    cams = make_cams(sc.K, kinv32(sc.K), R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    d_ref = engine.to_device(sc.depth_gt[ref]).clone()
    d_ref[40:80, 60:120] *= 1.2                                   # corrupted block: must be rejected
    d_src = [engine.to_device(sc.depth_gt[i]) for i in src]
    got = engine.geo_consistency(d_ref, d_src, cams, 1.0, 0.01)
    from oracle import oracle as orc
    ocams = orc.Cams(sc.K, kinv32(sc.K), R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    definition = orc.geo_consistency(d_ref.cpu().numpy(), [x.cpu().numpy() for x in d_src], ocams, 1.0, 0.01)
    assert np.array_equal(got.cpu().numpy(), definition)            # bit-exact against the written definition
    exp = torch_geo(d_ref, d_src, sc.K, R[ref], t[ref], [R[i] for i in src], [t[i] for i in src], 1.0, 0.01)
    mism = (got != exp).float().mean().item()
    assert mism < 5e-3, mism                                      # independent formulation: borderline pixels only
    g = got.cpu().numpy()
    assert g[40:80, 60:120].max() == 0
    inner = g[10:30, 60:140]
    assert (inner >= 3).mean() > 0.95                             # exact depths agree in (almost) every source
    assert set(np.unique(g)) <= {0.0, 1.0, 2.0, 3.0, 4.0}


def test_geometric_mode_filters_more_than_photometric():
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    sc = make_scene('wavy', 6, 128, 96, seed=21)
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
    sparse = np.stack([np.zeros(20), np.zeros(20), np.linspace(4.0, 6.0, 20)], axis=1)
    n = {}
    for mode in ('photometric', 'geometric'):
        pm = PatchMatchMVS(cam, scale=1.0, patch_size=7, num_iterations=2, num_samples=2, min_views=2, seed=5,
                           consistency=mode, geo_min_views=1, geo_rel_thresh=0.02, geo_px_thresh=2.0)
        pm.device_postfilter = False
        pm._filter_points = lambda p, c: (p, c)                    # compare raw fused clouds
        p, c = pm.reconstruct(sc.images, poses, sparse_points=sparse)
        n[mode] = len(p)
    assert 0 < n['geometric'] < n['photometric']


def test_geo_consistency_bitexact_with_skew_and_noise(engine):
    """Noisy depths (many borderline pixels), non-zero skew, zero / negative / NaN depths: kernel == definition."""
    from oracle import oracle as orc
    sc = make_scene('wavy', 6, 131, 97, seed=4)
    K = sc.K.copy()
    K[0, 1] = 0.7                                                 # skew: the sampling path honours it, so must this
    ref, src = 2, [1, 3, 0, 4, 5]
    R = [sc.poses[i][0] for i in range(6)]
    t = [sc.poses[i][1] for i in range(6)]
    args = (K, kinv32(K), R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    rng = np.random.default_rng(0)
    maps = [(sc.depth_gt[i] * (1 + 0.004 * rng.standard_normal(sc.depth_gt[i].shape))).astype(np.float32) for i in range(6)]
    maps[ref][5:9, :] = 0.0
    maps[ref][9:12, :] = -1.0
    maps[ref][12, :20] = np.nan
    maps[1][40:50, 40:90] = 0.0
    for px, rel in ((1.0, 0.01), (0.5, 0.002), (3.0, 0.05)):
        got = engine.geo_consistency(engine.to_device(maps[ref]), [engine.to_device(maps[i]) for i in src],
                                     make_cams(*args), px, rel).cpu().numpy()
        exp = orc.geo_consistency(maps[ref], [maps[i] for i in src], orc.Cams(*args), px, rel)
        assert np.array_equal(got, exp)
        assert 0 < (got > 0).mean() < 1
