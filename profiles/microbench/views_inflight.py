#!/usr/bin/env python
"""Throughput of N PatchMatch views in flight (the bench's timed step without fusion to points), for tuning runs:
    B200DENSE_LIB=<variant.so> python profiles/microbench/views_inflight.py [streams] [views]
Prints views/s and, from one serialised pass, the per-tag kernel durations.  (The B200D_NCC_CTAS / B200D_FUSE
environment switches it echoes belonged to the rejected persistent / fused variants of
profiles/experiments/r01_fused_persistent_mix.patch.txt; the shipped library ignores them.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200')]
import numpy as np
import torch

from b200dense.engine import DenseEngine, kinv32, make_cams

NS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
V = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = DenseEngine('cuda:0')
H, W, S, k = 1200, 1600, 5, 11
g = torch.Generator().manual_seed(0)
imgs = [torch.rand(H, W, generator=g).cuda() for _ in range(S + 2)]
K = np.array([[1440.0, 0, 800], [0, 1440.0, 600], [0, 0, 1]])


def pose(a):
    c, s = np.cos(a), np.sin(a)
    R = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
    return R, -R @ np.array([5 * s, 0, 5 - 5 * c])


P = [pose(np.radians(6.0 * (i - 3))) for i in range(S + 2)]
order = (2, 4, 1, 5, 0)
cams = make_cams(K, kinv32(K), P[3][0], P[3][1], [P[i][0] for i in order], [P[i][1] for i in order])
srcs = [imgs[i] for i in order]
outs = [dict(depth=torch.empty((H, W), device='cuda'), normal=torch.empty((H, W, 3), device='cuda'),
             confidence=torch.empty((H, W), device='cuda')) for _ in range(V)]
jobs = [(lambda slot, r=r: eng.pm_view(imgs[3], srcs, cams, k, 3, 8, 2.0, 10.0, seed=1, view_id=r, out=outs[r], slot=slot))
        for r in range(V)]
for _ in range(2):
    eng.run_views(jobs, NS)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    eng.run_views(jobs, NS)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
eng.profile_begin(8192)
eng.run_views(jobs[:2], 1)
torch.cuda.synchronize()
kms, tags = eng.profile_end()
per = {int(t): round(float(kms[tags == t].mean()), 4) for t in np.unique(tags)}
print(f"CTAS={os.environ.get('B200D_NCC_CTAS', '-')} FUSE={os.environ.get('B200D_FUSE', '-')} streams={NS}: "
      f"{ms:.2f} ms per {V} views = {V / ms * 1e3:.1f} views/s; serialised sum {float(kms.sum()) / 2:.3f} ms/view; {per}")
