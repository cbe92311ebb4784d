import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200'), os.path.join(ROOT, '3d-reconstruction-tool_b200', 'src')]
import numpy as np, torch
H, W = 1200, 1600
dev = 'cuda:0'
x = torch.rand(H, W).pin_memory()
xn = x.numpy()
t = torch.from_numpy(xn)
print('from_numpy(pinned.numpy()).is_pinned():', t.is_pinned())
def tm(f, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('h2d pinned tensor   ms', tm(lambda: x.to(dev, non_blocking=True)))
print('h2d from_numpy view ms', tm(lambda: torch.from_numpy(xn).to(dev, non_blocking=True)))
print('h2d from_numpy ascontig ms', tm(lambda: torch.from_numpy(np.ascontiguousarray(xn)).to(dev, dtype=torch.float32, non_blocking=True).contiguous()))
d = torch.rand(H, W, 3, device=dev)
def d2h():
    h = torch.empty(d.shape, dtype=d.dtype, pin_memory=True); h.copy_(d, non_blocking=True); return h
print('d2h pinned alloc+copy ms (first)', tm(d2h, 1)); print('d2h again', tm(d2h, 5))
keep = []
def d2h_keep():
    keep.append(d2h())
print('d2h keep (new alloc each) ms', tm(d2h_keep, 8))
keep.clear()
print('d2h keep after clear ms', tm(d2h_keep, 8))
s = torch.cuda.Stream()
def d2h_stream():
    with torch.cuda.stream(s):
        keep.append(d2h())
keep.clear()
print('d2h on side stream ms', tm(d2h_stream, 8))
