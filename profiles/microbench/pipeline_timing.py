#!/usr/bin/env python
"""Wall-clock split of the public `reconstruct()` pipelines at benchmark resolution (not the driver's bench):
prepare -> depth maps -> fusion -> post filter -> PLY, through the reference-facing classes."""
import os, sys, time, contextlib, io, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200'), os.path.join(ROOT, '3d-reconstruction-tool_b200', 'src')]
import numpy as np, torch
from b200dense.scenes import make_scene
from core.camera import Camera, CameraPose
from core.dense_stereo import DenseStereoReconstructor
from core.mvs_patchmatch import PatchMatchMVS
from core.utils import save_ply

V, W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1600, 1200
sc = make_scene('turntable', V, W, H, arc_deg=360.0 / 64, seed=0, device='cuda')
cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
sparse = np.random.default_rng(0).normal(size=(2000, 3)) * 1.2

def timed(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print(f'{label:58s} {time.perf_counter() - t0:8.3f} s', file=sys.stderr); return out

for rep in range(2):
    pm = PatchMatchMVS(cam, scale=1.0, num_iterations=3, min_views=3, num_src_views=5, seed=1)
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        p, c = timed(f'[rep {rep}] PatchMatchMVS.reconstruct {V} views {W}x{H}', lambda: pm.reconstruct(sc.images, poses, sparse_points=sparse))
    print('   ', [l for l in buf.getvalue().splitlines() if 'points' in l or 'filtering' in l][:3], file=sys.stderr)
    print(f'    -> {len(p):,} points', file=sys.stderr)
st = DenseStereoReconstructor(cam, scale=1.0, num_depths=64, patch_size=5)
with contextlib.redirect_stdout(io.StringIO()):
    ps, cs = timed(f'DenseStereoReconstructor.reconstruct {V} views D=64 k=5', lambda: st.reconstruct(sc.images, poses, max_pairs=30))
print(f'    -> {len(ps):,} points', file=sys.stderr)
with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()):
    big_p = np.tile(p, (max(1, 2_000_000 // max(len(p), 1)), 1)) if len(p) else np.zeros((0, 3))
    big_c = np.tile(c, (max(1, 2_000_000 // max(len(p), 1)), 1)) if len(p) else np.zeros((0, 3), np.uint8)
    timed(f'save_ply {len(big_p):,} points (native writer)', lambda: save_ply(big_p, big_c, os.path.join(d, 'x.ply')))
