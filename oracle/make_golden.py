#!/usr/bin/env python
"""Generate golden fixtures by EXECUTING the unmodified reference on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where the reference is
mounted read-only at /root/reference; the GPU box has no reference tree, so the
outputs are committed under tests/golden/ together with this script.

The reference has no tests or golden vectors of its own, so its
behaviour is pinned operationally: these fixtures are what
`core.dense_stereo.DenseStereoReconstructor` and `core.mvs_patchmatch.PatchMatchMVS`
(reference files src/core/dense_stereo.py, src/core/mvs_patchmatch.py) return with
torch 2.11.0 on CPU, fp32, default flags, on the synthetic scenes of
`b200dense.scenes`.  Nothing in the reference is edited; intermediate tensors are
captured by wrapping bound methods / torch.rand / torch.randn from the outside.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get('REF_SRC', '/root/reference/src')
sys.path.insert(0, REF_SRC)
sys.path.insert(0, os.path.join(ROOT, '3d-reconstruction-tool_b200'))

from core.camera import Camera, CameraPose                      # noqa: E402  (reference)
from core.dense_stereo import DenseStereoReconstructor          # noqa: E402  (reference)
from core.mvs_patchmatch import PatchMatchMVS                   # noqa: E402  (reference)
import core as _refcore                                         # noqa: E402
assert os.path.realpath(_refcore.__file__).startswith(os.path.realpath(REF_SRC)), _refcore.__file__

from b200dense.scenes import make_scene                         # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def scene_arrays(sc):
    n = len(sc.images)
    return dict(
        K=sc.K,
        R=np.stack([sc.poses[i][0] for i in range(n)]),
        t=np.stack([sc.poses[i][1] for i in range(n)]),
        bgr=np.stack([im['image'] for im in sc.images]),
    )


def ref_objects(sc):
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    poses = {i: CameraPose(R=sc.poses[i][0].copy(), t=sc.poses[i][1].copy()) for i in sc.poses}
    return cam, poses


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# --------------------------------------------------------------------------- plane sweep
def golden_plane_sweep(name, kind, n_views, W, H, D, k, ref_idx, dump_ncc, arc=8.0, thresh=0.8):
    sc = make_scene(kind, n_views, W, H, arc_deg=arc, seed=1)
    cam, poses = ref_objects(sc)
    obj = DenseStereoReconstructor(cam, scale=1.0, num_depths=D, patch_size=k, consistency_thresh=thresh)
    obj.device = torch.device('cpu')
    idxs = sorted(poses.keys())
    processed = obj._prepare_images(sc.images, idxs)
    nbrs = obj._find_neighbors(ref_idx, idxs, poses, k=6)
    dmin, dmax = sc.depth_range

    ncc_maps = []
    orig = obj._compute_ncc_torch

    def wrapped(a, b, p):
        out = orig(a, b, p)
        if dump_ncc:
            ncc_maps.append(out.detach().numpy().copy())
        return out
    obj._compute_ncc_torch = wrapped
    depth_map, conf, color = obj._compute_depth_map_gpu(ref_idx, nbrs, processed, poses, dmin, dmax)
    obj._compute_ncc_torch = orig

    depths = 1.0 / np.linspace(1 / dmax, 1 / dmin, D)
    pts, cols = obj._backproject(depth_map, conf, color, poses[ref_idx], min_confidence=obj.min_views - 0.5)
    K_inv32 = torch.inverse(torch.from_numpy(obj.K_scaled).float()).numpy()
    d = scene_arrays(sc)
    d.update(gray=np.stack([processed[i]['gray'] for i in idxs]), K_scaled=obj.K_scaled, K_inv32=K_inv32,
             depths=depths, neighbors=np.array(nbrs), ref_idx=ref_idx, patch=k, thresh=thresh,
             depth_map=depth_map, confidence=conf, points=pts, colors=np.ascontiguousarray(cols),
             depth_min=dmin, depth_max=dmax, min_confidence=obj.min_views - 0.5)
    if dump_ncc:
        d['ncc'] = np.stack(ncc_maps).reshape(D, len(nbrs), H, W)
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, 'pts', len(pts), 'conf hist', np.bincount(conf.astype(int).ravel()))


# --------------------------------------------------------------------------- PatchMatch
def golden_patchmatch(name, kind, n_views, W, H, k, iters, samples, ref_idx, n_src, seed,
                      store_rng, store_costs, arc=8.0):
    sc = make_scene(kind, n_views, W, H, arc_deg=arc, seed=2)
    cam, poses = ref_objects(sc)
    obj = PatchMatchMVS(cam, scale=1.0, patch_size=k, num_iterations=iters, num_samples=samples)
    obj.device = torch.device('cpu')
    obj.depth_min, obj.depth_max = sc.depth_range
    idxs = sorted(poses.keys())
    proc = obj._prepare_images(sc.images, idxs)
    srcs = obj._select_source_views(ref_idx, idxs, poses, k=n_src)

    rng, costs = [], []
    o_rand, o_randn, o_cost = torch.rand, torch.randn, obj._compute_patch_cost

    def w_rand(*a, **kw):
        out = o_rand(*a, **kw); rng.append(out.numpy().copy()); return out

    def w_randn(*a, **kw):
        out = o_randn(*a, **kw); rng.append(out.numpy().copy()); return out

    def w_cost(*a, **kw):
        out = o_cost(*a, **kw); costs.append(out.numpy().copy()); return out

    torch.manual_seed(seed)
    torch.rand, torch.randn, obj._compute_patch_cost = w_rand, w_randn, w_cost
    try:
        dn = obj._patchmatch_cuda(ref_idx, srcs, proc, poses)
    finally:
        torch.rand, torch.randn = o_rand, o_randn
        obj._compute_patch_cost = o_cost

    assert len(rng) == 3 + iters * samples * 2
    # RNG call order (mvs_patchmatch.py:271,279,280,471,475): U0, Nx, Ny, then per sample U (H,W), N (H,W,3)
    init_u, init_nx, init_ny = rng[0], rng[1], rng[2]
    ref_u = np.stack(rng[3::2]) if iters * samples else np.zeros((0, H, W), np.float32)
    ref_n = np.stack(rng[4::2]) if iters * samples else np.zeros((0, H, W, 3), np.float32)

    # initial hypotheses exactly as the reference computed them (replay of lines 268-281 on the captured RNG)
    import torch.nn.functional as F
    ld0, ld1 = np.log(obj.depth_min), np.log(obj.depth_max)
    depth0 = torch.exp(torch.from_numpy(init_u) * (ld1 - ld0) + ld0).numpy()
    n0 = torch.zeros(H, W, 3)
    n0[:, :, 2] = -1
    n0[:, :, 0] = torch.from_numpy(init_nx) * 0.3
    n0[:, :, 1] = torch.from_numpy(init_ny) * 0.3
    normal0 = F.normalize(n0, dim=-1).numpy()

    dm = {ref_idx: dn}
    pts, cols = obj._fuse_depth_maps(dm, proc, poses)
    K_inv32 = torch.inverse(torch.from_numpy(obj.K_scaled.astype(np.float32))).numpy()
    d = scene_arrays(sc)
    d.update(gray=np.stack([proc[i]['gray'] for i in idxs]), K_scaled=obj.K_scaled, K_inv32=K_inv32,
             ref_idx=ref_idx, sources=np.array(srcs), patch=k, iters=iters, samples=samples, seed=seed,
             depth_min=obj.depth_min, depth_max=obj.depth_max, min_views=obj.min_views,
             depth=dn.depth, normal=dn.normal, confidence=dn.confidence,
             points=pts, colors=np.ascontiguousarray(cols).astype(np.uint8),
             rng_sha=np.array([sha(init_u), sha(init_nx), sha(init_ny), sha(ref_u), sha(ref_n)]))
    if store_rng:
        d.update(depth0=depth0, normal0=normal0, ref_u=ref_u, ref_n=ref_n,
                 init_u=init_u, init_nx=init_nx, init_ny=init_ny)
    if store_costs:
        d['costs'] = np.stack(costs)
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, 'srcs', srcs, 'conf hist', np.bincount(dn.confidence.astype(int).ravel()), 'pts', len(pts),
          'finite cost frac', np.isfinite(costs[-1]).mean())


# --------------------------------------------------------------------------- host heuristics + full reconstruct
def golden_host(name):
    out = {}
    for tag, kind, n, arc in (('plane7', 'plane', 7, 8.0), ('tt12', 'turntable', 12, 12.0), ('tt40', 'turntable', 40, 9.0)):
        sc = make_scene(kind, n, 32, 24, arc_deg=arc, seed=3)
        cam, poses = ref_objects(sc)
        pm = PatchMatchMVS(cam, scale=1.0)
        st = DenseStereoReconstructor(cam, scale=1.0)
        idxs = sorted(poses.keys())
        out[f'{tag}_R'] = np.stack([sc.poses[i][0] for i in idxs])
        out[f'{tag}_t'] = np.stack([sc.poses[i][1] for i in idxs])
        for k in (4, 5):
            out[f'{tag}_src_k{k}'] = np.array([pm._select_source_views(r, idxs, poses, k=k) + [-1] * (k - len(pm._select_source_views(r, idxs, poses, k=k))) for r in idxs])
        out[f'{tag}_nbr_k6'] = np.array([(st._find_neighbors(r, idxs, poses, k=6) + [-1] * 6)[:6] for r in idxs])
        pm._estimate_depth_range(poses, None)
        out[f'{tag}_range_cams'] = np.array([pm.depth_min, pm.depth_max])
        rs = np.random.default_rng(5).normal(size=(500, 3)) * 1.5
        pm._estimate_depth_range(poses, rs)
        out[f'{tag}_sparse'] = rs
        out[f'{tag}_range_sparse'] = np.array([pm.depth_min, pm.depth_max])
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, 'ok')


def golden_reconstruct(name):
    """Full public-API runs: `reconstruct()` of both classes incl. post filters (scale 1.0 and 0.5)."""
    out = {}
    sc = make_scene('wavy', 5, 96, 72, arc_deg=8.0, seed=4)
    cam, poses = ref_objects(sc)
    out.update({f'sc_{k}': v for k, v in scene_arrays(sc).items()})
    for scale in (1.0, 0.5):
        st = DenseStereoReconstructor(cam, scale=scale, num_depths=24, patch_size=5)
        st.device = torch.device('cpu')
        p, c = st.reconstruct(sc.images, poses, max_pairs=30)
        out[f'stereo_s{scale}_points'] = p
        out[f'stereo_s{scale}_colors'] = np.ascontiguousarray(c)
        pm = PatchMatchMVS(cam, scale=scale, patch_size=7, num_iterations=2, num_samples=3, min_views=2)
        pm.device = torch.device('cpu')
        torch.manual_seed(11)
        sparse = np.stack([np.zeros(50), np.zeros(50), np.linspace(4.0, 6.0, 50)], axis=1)
        p, c = pm.reconstruct(sc.images, poses, sparse_points=sparse)
        out[f'mvs_s{scale}_points'] = p
        out[f'mvs_s{scale}_colors'] = np.ascontiguousarray(c)
        out[f'mvs_s{scale}_range'] = np.array([pm.depth_min, pm.depth_max])
        out['sparse'] = sparse
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, {k: v.shape for k, v in out.items() if 'points' in k})


if __name__ == '__main__':
    torch.set_num_threads(8)
    golden_plane_sweep('planesweep_plane_64x48_k5.npz', 'plane', 5, 64, 48, D=8, k=5, ref_idx=2, dump_ncc=True)
    golden_plane_sweep('planesweep_wavy_80x60_k7.npz', 'wavy', 5, 80, 60, D=16, k=7, ref_idx=1, dump_ncc=True)
    golden_plane_sweep('planesweep_tt_161x97_k7.npz', 'turntable', 8, 161, 97, D=32, k=7, ref_idx=3, dump_ncc=False, arc=7.0)
    golden_patchmatch('patchmatch_wavy_64x48_k11.npz', 'wavy', 5, 64, 48, k=11, iters=3, samples=8, ref_idx=2,
                      n_src=4, seed=7, store_rng=True, store_costs=True)
    golden_patchmatch('patchmatch_plane_57x43_k5.npz', 'plane', 6, 57, 43, k=5, iters=2, samples=3, ref_idx=0,
                      n_src=5, seed=9, store_rng=True, store_costs=True)
    golden_patchmatch('patchmatch_tt_160x120_k11.npz', 'turntable', 9, 160, 120, k=11, iters=3, samples=8, ref_idx=4,
                      n_src=5, seed=123, store_rng=False, store_costs=False, arc=7.0)
    # the remaining supported window sizes (9, 13): pins the oracle's window sums / NCC against the reference there too
    golden_plane_sweep('planesweep_plane_60x44_k9.npz', 'plane', 5, 60, 44, D=8, k=9, ref_idx=2, dump_ncc=True)
    golden_plane_sweep('planesweep_wavy_72x54_k13.npz', 'wavy', 5, 72, 54, D=8, k=13, ref_idx=1, dump_ncc=True)
    golden_patchmatch('patchmatch_wavy_60x45_k9.npz', 'wavy', 5, 60, 45, k=9, iters=2, samples=3, ref_idx=2,
                      n_src=4, seed=21, store_rng=True, store_costs=True)
    golden_patchmatch('patchmatch_plane_64x48_k13.npz', 'plane', 6, 64, 48, k=13, iters=2, samples=2, ref_idx=3,
                      n_src=5, seed=22, store_rng=True, store_costs=True)
    golden_host('host_heuristics.npz')
    golden_reconstruct('reconstruct_wavy_96x72.npz')
