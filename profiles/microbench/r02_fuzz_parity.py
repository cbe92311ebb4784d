"""Randomised parity stress (GPU box): for `seconds` of wall time draw random sizes / window sizes / neighbour counts /
thresholds / image transforms and check
  * plane sweep: B200D_PARITY (bounded decision) == B200D_PARITY_CHAINS == oracle (counts, arg-max, depth),
  * PatchMatch cost + confidence of a random depth map == oracle (bit-exact, NaN-aware); at small sizes a full
    PatchMatch view with injected random numbers == oracle (depth, normal, confidence, best cost),
  * geometric-consistency and view-propagation proposal kernels == their definitions in oracle/dense_oracle.c.
Prints one summary line; exits non-zero on the first mismatch.   python profiles/microbench/r02_fuzz_parity.py [seconds]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from b200dense import PARITY, PARITY_CHAINS
from b200dense.engine import DenseEngine, kinv32, make_cams
from b200dense.scenes import gray_f32, make_scene
from oracle import oracle as orc

F = np.float32
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
eng = DenseEngine('cuda:0')
rng = np.random.default_rng(int(os.environ.get('FUZZ_SEED', '0')))
t_end = time.time() + seconds
n_cases = {'sweep': 0, 'pm': 0, 'pm_view': 0, 'geo': 0, 'vprop': 0}


def transform(g, how):
    if how == 'scale':
        return (g * F(rng.choice([1e-3, 0.1, 3.0, 50.0]))).astype(F)
    if how == 'offset':
        return (g + F(rng.choice([-0.5, 0.25, 2.0]))).astype(F)
    if how == 'quant':
        q = int(rng.choice([2, 4, 16]))
        return (np.floor(g * q) / q).astype(F)
    if how == 'flatband':
        g = g.copy()
        a = int(rng.integers(0, g.shape[1] // 2))
        g[:, a:a + g.shape[1] // 3] = F(rng.random())
        return g
    return g


while time.time() < t_end:
    W, H = int(rng.integers(40, 420)), int(rng.integers(36, 300))
    if rng.random() < 0.4:
        W, H = int(rng.integers(40, 160)), int(rng.integers(36, 120))      # small cases also run a full PatchMatch view
    n = int(rng.integers(3, 10))
    kind = str(rng.choice(['wavy', 'plane', 'turntable']))
    sc = make_scene(kind, n, W, H, seed=int(rng.integers(0, 1 << 30)), arc_deg=float(rng.uniform(2.0, 9.0)))
    gray = gray_f32(sc)
    how = str(rng.choice(['none', 'scale', 'offset', 'quant', 'flatband']))
    gray = [transform(g, how) for g in gray]
    ref = int(rng.integers(0, n))
    others = [i for i in range(n) if i != ref]
    rng.shuffle(others)
    nb = others[:int(rng.integers(1, min(8, len(others)) + 1))]
    Ki = kinv32(sc.K)
    args = (sc.K, Ki, sc.poses[ref][0], sc.poses[ref][1], [sc.poses[i][0] for i in nb], [sc.poses[i][1] for i in nb])
    ecams, ocams = make_cams(*args), orc.Cams(*args)
    k = int(rng.choice([3, 5, 7, 9, 11, 13]))
    dmin, dmax = sc.depth_range
    d = eng.to_device
    dref, dnb = d(gray[ref]), [d(gray[i]) for i in nb]
    # ---- plane sweep
    D = int(rng.integers(3, 20))
    depths = (1.0 / np.linspace(1 / dmax, 1 / dmin, D)).astype(F)
    th = float(rng.choice([0.3, 0.5, 0.8, 0.9, 0.97]))
    hy = eng.plane_sweep(dref, dnb, ecams, depths, k, th, arith=PARITY)
    ch = eng.plane_sweep(dref, dnb, ecams, depths, k, th, arith=PARITY_CHAINS)
    o = orc.plane_sweep(gray[ref], [gray[i] for i in nb], ocams, depths, k, th)
    for key in ('best_idx', 'confidence', 'depth_map'):
        a, b = hy[key].cpu().numpy(), ch[key].cpu().numpy()
        if not (np.array_equal(a, b) and np.array_equal(a, o[key])):
            print(f'MISMATCH sweep {key}: {kind} {W}x{H} k={k} N={len(nb)} D={D} th={th} {how}: '
                  f'hybrid!=chains {int((a != b).sum())}, hybrid!=oracle {int((a != o[key]).sum())}')
            sys.exit(1)
    n_cases['sweep'] += 1
    # ---- PatchMatch cost / confidence of a random depth map
    depth = (rng.random((H, W), dtype=F) * F(dmax - dmin) + F(dmin))
    depth[: H // 3] = sc.depth_gt[ref][: H // 3]
    for mode in (0, 1):
        got = eng.pm_cost(dref, dnb, ecams, d(depth), k, mode).cpu().numpy()
        exp = orc.pm_cost(gray[ref], [gray[i] for i in nb], ocams, depth, k, mode)
        if not np.array_equal(got, exp, equal_nan=True):
            print(f'MISMATCH pm_cost mode {mode}: {kind} {W}x{H} k={k} S={len(nb)} {how}: {int((got != exp).sum())}')
            sys.exit(1)
    n_cases['pm'] += 1
    # ---- a full PatchMatch view with injected random numbers (small sizes: the oracle runs every evaluation on the CPU)
    if W * H <= 160 * 120 and len(nb) >= 2:
        iters, samples = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        n0 = rng.standard_normal((H, W, 3)).astype(F)
        n0 /= np.linalg.norm(n0, axis=-1, keepdims=True)
        ru = rng.random((iters * samples, H, W), dtype=F)
        rn = rng.standard_normal((iters * samples, H, W, 3)).astype(F)
        out = eng.pm_view(dref, dnb, ecams, k, iters, samples, dmin, dmax, init_depth=d(depth), init_normal=d(n0),
                          rand_u=d(ru), rand_n=d(rn), want_best=True)
        exp = orc.pm_view(gray[ref], [gray[i] for i in nb], ocams, k, iters, samples, dmin, dmax, depth, n0, ru, rn)
        for key in ('depth', 'normal', 'confidence', 'best_cost'):
            if not np.array_equal(out[key].cpu().numpy(), exp[key], equal_nan=True):
                print(f'MISMATCH pm_view {key}: {kind} {W}x{H} k={k} S={len(nb)} it={iters} sm={samples} {how}')
                sys.exit(1)
        n_cases['pm_view'] += 1
    # ---- cross-view extension kernels vs their definitions
    maps = [(sc.depth_gt[i] * (1 + 0.01 * rng.standard_normal(sc.depth_gt[i].shape))).astype(F) for i in range(n)]
    px, rel = float(rng.choice([0.5, 1.0, 2.5])), float(rng.choice([0.003, 0.01, 0.05]))
    got = eng.geo_consistency(d(maps[ref]), [d(maps[i]) for i in nb], ecams, px, rel).cpu().numpy()
    if not np.array_equal(got, orc.geo_consistency(maps[ref], [maps[i] for i in nb], ocams, px, rel)):
        print(f'MISMATCH geo: {kind} {W}x{H} S={len(nb)}')
        sys.exit(1)
    n_cases['geo'] += 1
    nrm = rng.standard_normal((n, H, W, 3)).astype(F)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    state = dict(depth=d(maps[ref]).clone(), normal=d(nrm[ref]).clone(), best_cost=torch.full((H, W), float('inf'), device=eng.device))
    costs = eng.pm_viewprop(dref, dnb, ecams, k, dmin, dmax, [d(maps[i]) for i in nb], [d(nrm[i]) for i in nb], state,
                            dump_costs=True).cpu().numpy()
    dep, nor, best = maps[ref].copy(), nrm[ref].copy(), np.full((H, W), np.inf, F)
    for s in range(len(nb)):
        cd, cn, allow = orc.viewprop_candidates(dep, nor, maps[nb[s]], nrm[nb[s]], s, ocams, dmin, dmax)
        c = orc.pm_cost(gray[ref], [gray[i] for i in nb], ocams, cd, k, 0)
        if not np.array_equal(costs[s], c, equal_nan=True):
            print(f'MISMATCH viewprop cost source {s}: {kind} {W}x{H} k={k}')
            sys.exit(1)
        win = (c < best) & allow
        dep[win], nor[win], best[win] = cd[win], cn[win], c[win]
    if not (np.array_equal(state['depth'].cpu().numpy(), dep) and np.array_equal(state['normal'].cpu().numpy(), nor)
            and np.array_equal(state['best_cost'].cpu().numpy(), best, equal_nan=True)):
        print(f'MISMATCH viewprop state: {kind} {W}x{H} k={k}')
        sys.exit(1)
    n_cases['vprop'] += 1
print(f'fuzz ok: {n_cases} random cases in {seconds:.0f} s, 0 mismatches (seed {os.environ.get("FUZZ_SEED", "0")})')
