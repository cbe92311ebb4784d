#!/usr/bin/env python
"""Per-kernel timing of tuning variants of libb200dense.so (experiments; not part of the product).

    python profiles/microbench/time_kernels.py variant1.so variant2.so ...

For each library: PatchMatch of one 1600x1200 view (S=5, k=11) twice, per-launch CUDA-event durations
from b200d_profile_begin/_end, and one 1920x1080 D=64 N=6 k=7 plane sweep."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200')]
import numpy as np
import torch

from b200dense import _cabi


def run(libpath):
    _cabi._lib = None
    _cabi.LIB_PATH = libpath
    from b200dense.engine import DenseEngine, kinv32, make_cams
    eng = DenseEngine('cuda:0')
    H, W, S, k = 1200, 1600, 5, 11
    g = torch.Generator().manual_seed(0)
    imgs = [torch.rand(H, W, generator=g).cuda() for _ in range(S + 2)]
    K = np.array([[1440.0, 0, 800], [0, 1440.0, 600], [0, 0, 1]])
    def pose(a):
        c, s = np.cos(a), np.sin(a)
        R = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
        C = np.array([5 * s, 0, 5 - 5 * c])
        return R, -R @ C
    P = [pose(np.radians(6.0 * (i - 3))) for i in range(S + 2)]
    cams = make_cams(K, kinv32(K), P[3][0], P[3][1], [P[i][0] for i in (2, 4, 1, 5, 0)], [P[i][1] for i in (2, 4, 1, 5, 0)])
    srcs = [imgs[i] for i in (2, 4, 1, 5, 0)]
    for _ in range(2):
        eng.pm_view(imgs[3], srcs, cams, k, 3, 8, 2.0, 10.0, seed=1)
    torch.cuda.synchronize()
    eng.profile_begin(4096)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        eng.pm_view(imgs[3], srcs, cams, k, 3, 8, 2.0, 10.0, seed=1)
    e1.record()
    torch.cuda.synchronize()
    ms, tags = eng.profile_end()
    out = {int(t): float(ms[tags == t].mean()) for t in np.unique(tags)}
    view_ms = e0.elapsed_time(e1) / 2
    # plane sweep
    Hs, Ws = 1080, 1920
    im2 = [torch.rand(Hs, Ws, generator=g).cuda() for _ in range(7)]
    K2 = np.array([[1728.0, 0, 960], [0, 1728.0, 540], [0, 0, 1]])
    cams2 = make_cams(K2, kinv32(K2), P[3][0], P[3][1], [P[i][0] for i in (2, 4, 1, 5, 0, 6)], [P[i][1] for i in (2, 4, 1, 5, 0, 6)])
    depths = (1.0 / np.linspace(1 / 10.0, 1 / 2.0, 64)).astype(np.float32)
    nb = [im2[i] for i in (2, 4, 1, 5, 0, 6)]
    eng.plane_sweep(im2[3], nb, cams2, depths, 7, 0.8)
    torch.cuda.synchronize()
    eng.profile_begin(4096)
    e0.record()
    eng.plane_sweep(im2[3], nb, cams2, depths, 7, 0.8)
    e1.record()
    torch.cuda.synchronize()
    ms2, tags2 = eng.profile_end()
    sw = e0.elapsed_time(e1)
    out2 = {int(t): (float(ms2[tags2 == t].mean()), int((tags2 == t).sum())) for t in np.unique(tags2)}
    print(f"{os.path.basename(libpath):40s} view {view_ms:7.3f} ms   sweep(D=64) {sw:7.2f} ms = {Hs * Ws * 64 / sw / 1e6:.2f} Gpix*planes/s")
    print('   pm tags (ms):', {t: round(v, 4) for t, v in out.items()})
    print('   sweep tags (ms, n):', {t: (round(v, 4), c) for t, (v, c) in out2.items()})


if __name__ == '__main__':
    for p in sys.argv[1:]:
        run(os.path.abspath(p))
