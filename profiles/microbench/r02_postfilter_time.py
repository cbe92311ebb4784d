"""Phase timing of b200dense.postfilter.filter_points on the GPU (synthetic cloud; sizes like the fused clouds of the
full4k benchmark).  python profiles/microbench/r02_postfilter_time.py [n_points ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, '3d-reconstruction-tool_b200'))
import torch
from b200dense import postfilter as pf
from b200dense.engine import DenseEngine

eng = DenseEngine('cuda:0')


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for n in [int(x) for x in (sys.argv[1:] or ['600000', '6400000', '60000000'])]:
    g = torch.Generator(device='cuda').manual_seed(0)
    pts = (torch.randn((n, 3), generator=g, device='cuda', dtype=torch.float64) * 1.5)
    col = torch.randint(0, 255, (n, 3), generator=g, device='cuda', dtype=torch.uint8)
    for rep in range(3):
        t0 = sync()
        c = pf._median_axis0(pts)
        t1 = sync()
        d = pts - c
        dist = torch.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        thr = pf._percentile_linear(torch.sort(dist).values, 95)
        t2 = sync()
        keep = dist < thr
        p2, c2 = pts[keep], col[keep]
        t3 = sync()
        idx = pf._first_of_each_voxel(p2, 0.01, cast_int32=False)
        t4 = sync()
        fp, fc = p2[idx], c2[idx]
        t5 = sync()
        hp, hc = eng.to_host([fp, fc])
        t6 = sync()
        print(f'n={n:>9d} rep {rep}: median {1e3*(t1-t0):7.2f}  dist+sort+pct {1e3*(t2-t1):7.2f}  mask {1e3*(t3-t2):7.2f}  '
              f'voxel sort {1e3*(t4-t3):7.2f}  gather {1e3*(t5-t4):7.2f}  download {1e3*(t6-t5):7.2f}  '
              f'total {1e3*(t6-t0):7.2f} ms  -> {len(hp)} points')
        del hp, hc
