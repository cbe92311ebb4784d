"""Small driver for ncu captures of the separable-sum kernels: a few planes of the bounded plane sweep at 1920x1080 (N=6,
k=7) and one B200D_FAST PatchMatch view at 1600x1200 (S=5, k=11).  Usage (on the GPU box):
    ncu --set full --clock-control none --import-source on -k regex:'ncc_sep_kernel|sample_fast' --launch-skip 8 -c 6 \
        -o gpurun_out/r02_sep python profiles/microbench/r02_sep_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200'), os.path.join(ROOT, '3d-reconstruction-tool_b200', 'src')):
    sys.path.insert(0, p)
import numpy as np
import torch
from b200dense import FAST, PARITY
from b200dense.engine import DenseEngine, kinv32, make_cams
from b200dense.scenes import gray_f32, make_scene

mode = sys.argv[1] if len(sys.argv) > 1 else 'both'
eng = DenseEngine('cuda:0')
if mode in ('both', 'sweep'):
    W, H, D, k = 1920, 1080, 12, 7
    sc = make_scene('turntable', 64, W, H, arc_deg=360.0 / 64, seed=0, device='cuda', render=list(range(8)))
    g = gray_f32(sc)
    r, nb = 3, [2, 4, 1, 5, 0, 6]
    cams = make_cams(sc.K, kinv32(sc.K), sc.poses[r][0], sc.poses[r][1], [sc.poses[i][0] for i in nb], [sc.poses[i][1] for i in nb])
    depths = (1.0 / np.linspace(1 / sc.depth_range[1], 1 / sc.depth_range[0], 256)).astype(np.float32)[100:100 + D]
    d = eng.to_device
    eng.plane_sweep(d(g[r]), [d(g[i]) for i in nb], cams, depths, k, 0.8, arith=PARITY)
    torch.cuda.synchronize()
if mode in ('both', 'pm'):
    W, H = 1600, 1200
    sc = make_scene('turntable', 64, W, H, arc_deg=360.0 / 64, seed=0, device='cuda', render=list(range(12)))
    g = gray_f32(sc)
    r, src = 5, [3, 7, 2, 8, 1]
    cams = make_cams(sc.K, kinv32(sc.K), sc.poses[r][0], sc.poses[r][1], [sc.poses[i][0] for i in src], [sc.poses[i][1] for i in src])
    d = eng.to_device
    arith = FAST if (len(sys.argv) < 3 or sys.argv[2] == 'fast') else PARITY
    eng.pm_view(d(g[r]), [d(g[i]) for i in src], cams, 11, 1, 4, sc.depth_range[0], sc.depth_range[1], seed=1, view_id=r, arith=arith)
    torch.cuda.synchronize()
print('probe done')
