#!/usr/bin/env python
"""Wall-clock split of PatchMatchMVS.reconstruct() (public API, 1600x1200): where the time outside the kernels goes."""
import os, sys, time, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200'), os.path.join(ROOT, '3d-reconstruction-tool_b200', 'src')]
import numpy as np, torch
from b200dense.scenes import make_scene
from core.camera import Camera, CameraPose
from core.mvs_patchmatch import PatchMatchMVS

V = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sc = make_scene('turntable', V, 1600, 1200, arc_deg=360.0 / 64, seed=0, device='cuda')
cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
sparse = np.random.default_rng(0).normal(size=(2000, 3)) * 1.2
pm = PatchMatchMVS(cam, scale=1.0, num_iterations=3, min_views=3, num_src_views=5, seed=1)

def tick(label, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); print(f'  {label:34s} {1e3 * (t - t0):8.1f} ms'); return t

for rep in range(3):
    print(f'rep {rep}')
    with contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p, c = pm.reconstruct(sc.images, poses, sparse_points=sparse)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f'  reconstruct() total                {1e3 * (t1 - t0):8.1f} ms   -> {len(p):,} points')
    # the same phases by hand
    t = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        pm._estimate_depth_range(poses, sparse)
    t = tick('_estimate_depth_range', t)
    idx = sorted(poses)
    proc = pm._prepare_images_device(sc.images, idx)
    t = tick('_prepare_images_device', t)
    eng = pm.engine
    jobs = [(lambda slot, r=r: pm._patchmatch_device(r, pm._select_source_views(r, idx, poses, k=5), proc, poses, slot=slot)) for r in idx]
    res = eng.run_views(jobs, 4)
    t = tick('depth maps (run_views)', t)
    Kinv = np.linalg.inv(pm.K_scaled)
    local = {}
    for r, out in zip(idx, res):
        pts, cols = eng.mask_backproject(out['depth'], out['confidence'], pm._color_on_device(proc, r), Kinv, poses[r].R, poses[r].t, 3.0, 0)
        local[r] = (pts.clone(), cols.clone())
    t = tick('mask + back-projection', t)
    from b200dense import dist as bdist
    pts_d, col_d = bdist.gather_point_clouds(idx, local, eng.device)
    t = tick('gather', t)
    from b200dense.postfilter import filter_points
    fp, fc = filter_points(pts_d, col_d)
    t = tick('filter_points', t)
    a, b = fp.cpu().numpy(), fc.cpu().numpy()
    t = tick('download', t)
