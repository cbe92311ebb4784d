import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200')]
import numpy as np, torch
from b200dense.engine import DenseEngine
eng = DenseEngine('cuda:0')
x = torch.rand(48, 64, device='cuda')
m, v = eng.ref_stats(x, 5)
torch.cuda.synchronize()
print('ok', m.mean().item(), v.mean().item())
