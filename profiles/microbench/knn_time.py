import sys, time, numpy as np, torch
sys.path[:0]=['/root/repo','/root/repo/3d-reconstruction-tool_b200']
from b200dense.engine import DenseEngine
eng=DenseEngine('cuda:0')
for n in (100000, 500000):
    pts=torch.from_numpy(np.random.default_rng(0).normal(size=(n,3))).cuda()
    eng.knn_mean_dist(pts,20); torch.cuda.synchronize()
    t=time.perf_counter(); out=eng.knn_mean_dist(pts,20); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(n, f'{dt*1e3:.1f} ms', f'{n*n/dt/1e12:.2f} T pairs/s')
