import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200'), os.path.join(ROOT, '3d-reconstruction-tool_b200', 'src')]
import numpy as np, torch, contextlib, io
from core.camera import Camera, CameraPose
from core.mvs_patchmatch import PatchMatchMVS
H, W = 1200, 1600
K = np.array([[1440.0, 0, 800], [0, 1440.0, 600], [0, 0, 1]])
def pose(a):
    c, s = np.cos(a), np.sin(a); R = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]]); C = np.array([5 * s, 0, 5 - 5 * c]); return CameraPose(R=R, t=-R @ C)
n = 12
poses = {i: pose(np.radians(6.0 * (i - 6))) for i in range(n)}
pm = PatchMatchMVS(Camera(K=K, dist=np.zeros(5)), scale=1.0, num_src_views=5)
pm.depth_min, pm.depth_max = 2.0, 10.0
pm.cache_images = False
pinned = {i: torch.rand(H, W).pin_memory() for i in range(n)}
proc = {i: {'gray': pinned[i].numpy(), 'shape': (H, W)} for i in range(n)}
refs = list(range(2, 10))
srcs = {r: [r - 1, r + 1, r - 2, r + 2, (r + 3) % n] for r in refs}
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    maps = pm.compute_depth_maps(proc, poses, refs, srcs)
    t1 = time.perf_counter()
    del maps
    print(f'step {it}: {1e3 * (t1 - t0):.1f} ms')
# breakdown
import torch.autograd.profiler as prof
t0 = time.perf_counter()
for i in range(n): pm._gray_on_device(proc, i)
torch.cuda.synchronize(); print('upload 12 grays ms', 1e3 * (time.perf_counter() - t0))
t0 = time.perf_counter(); hs = [torch.empty((H, W, 3), pin_memory=True) for _ in range(8)]; print('8 pinned allocs ms', 1e3 * (time.perf_counter() - t0))
