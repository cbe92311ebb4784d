"""GPU (-m gpu): the opt-in red-black propagation schedule (EXTENSION -- the reference has no checkerboard sweep,
SURVEY F2, so there is no dackey-wav oracle for it).  Its definition is restated here in NumPy on top of the oracle's
cost evaluation (the same `_compute_patch_cost` arithmetic the graded schedule is pinned with) and the kernel path must
reproduce it bit for bit; the reference schedule must be untouched by the new flag."""
import numpy as np
import pytest
import torch

from b200dense.engine import kinv32, make_cams
from b200dense.scenes import gray_f32, make_scene
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
F = np.float32


def _shift(a, dy, dx, fill):
    """neighbour[y,x] = a[y+dy, x+dx], `fill` outside (the border convention of mvs_patchmatch.py:429-442)."""
    H, W = a.shape[:2]
    out = np.full_like(a, fill)
    ys, yd = (slice(dy, H), slice(0, H - dy)) if dy >= 0 else (slice(0, H + dy), slice(-dy, H))
    xs, xd = (slice(dx, W), slice(0, W - dx)) if dx >= 0 else (slice(0, W + dx), slice(-dx, W))
    out[yd, xd] = a[ys, xs]
    return out


def _redblack_numpy(gray, ref, src, ocams, k, iters, samples, dmin, dmax, depth, normal, ru, rn):
    H, W = depth.shape
    depth, normal = depth.copy(), normal.copy()
    best = np.full((H, W), np.inf, F)
    yy, xx = np.mgrid[0:H, 0:W]
    cost_of = lambda d: orc.pm_cost(gray[ref], [gray[i] for i in src], ocams, np.ascontiguousarray(d), k, 0)
    for it in range(iters):
        for color in (0, 1):
            for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                cd = _shift(depth, dy, dx, F(dmin))
                cn = _shift(normal, dy, dx, F(0))
                cost = cost_of(cd)
                win = (cost < best) & (((xx + yy) & 1) == color)
                depth[win], normal[win], best[win] = cd[win], cn[win], cost[win]
        dr = F((dmax - dmin) * 0.5 ** it)
        nr = F(0.5 * 0.5 ** it)
        for sm in range(samples):
            u, r = ru[it * samples + sm], rn[it * samples + sm]
            cd = np.clip(depth + (u * F(2) - F(1)) * dr, F(dmin), F(dmax)).astype(F)
            c = normal + r * nr
            nn = np.sqrt((c[..., 0] * c[..., 0] + c[..., 1] * c[..., 1]) + c[..., 2] * c[..., 2])
            cn = c / np.maximum(nn, F(1e-12))[..., None]
            cost = cost_of(cd)
            win = cost < best
            depth[win], normal[win], best[win] = cd[win], cn[win], cost[win]
    conf = orc.pm_cost(gray[ref], [gray[i] for i in src], ocams, depth, k, 1)
    return depth, normal, conf, best


@pytest.mark.parametrize('shape', [(64, 48), (57, 43)])
def test_redblack_schedule_equals_its_definition(engine, shape):
    W, H = shape
    sc = make_scene('wavy', 5, W, H, seed=21, arc_deg=6.0)
    gray = gray_f32(sc)
    ref, src = 2, [1, 3, 0, 4]
    Ki = kinv32(sc.K)
    R = [sc.poses[i][0] for i in range(5)]
    t = [sc.poses[i][1] for i in range(5)]
    ecams = make_cams(sc.K, Ki, R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    ocams = orc.Cams(sc.K, Ki, R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    dmin, dmax = sc.depth_range
    k, iters, samples = 7, 2, 2
    g = np.random.default_rng(5)
    depth0 = (g.random((H, W), dtype=F) * F(dmax - dmin) + F(dmin))
    n0 = g.standard_normal((H, W, 3), dtype=F)
    n0 /= np.linalg.norm(n0, axis=-1, keepdims=True)
    ru = g.random((iters * samples, H, W), dtype=F)
    rn = g.standard_normal((iters * samples, H, W, 3), dtype=F)
    d = engine.to_device
    args = (d(gray[ref]), [d(gray[i]) for i in src], ecams, k, iters, samples, dmin, dmax)
    kw = dict(init_depth=d(depth0), init_normal=d(n0), rand_u=d(ru), rand_n=d(rn), want_best=True)
    out = engine.pm_view(*args, schedule=1, dump_costs=True, **kw)
    assert out['costs'].shape[0] == iters * (8 + samples)
    e_depth, e_normal, e_conf, e_best = _redblack_numpy(gray, ref, src, ocams, k, iters, samples, dmin, dmax, depth0, n0, ru, rn)
    assert np.array_equal(out['depth'].cpu().numpy(), e_depth)
    assert np.array_equal(out['normal'].cpu().numpy(), e_normal)
    assert np.array_equal(out['confidence'].cpu().numpy(), e_conf)
    assert np.array_equal(out['best_cost'].cpu().numpy(), e_best, equal_nan=True)
    # the flag leaves the graded schedule alone
    a = engine.pm_view(*args, schedule=0, **kw)
    b = orc.pm_view(gray[ref], [gray[i] for i in src], ocams, k, iters, samples, dmin, dmax, depth0, n0, ru, rn)
    assert np.array_equal(a['depth'].cpu().numpy(), b['depth'])
    assert np.array_equal(a['normal'].cpu().numpy(), b['normal'])
    # and red-black really is a different walk
    assert not np.array_equal(out['depth'].cpu().numpy(), a['depth'].cpu().numpy())


def test_redblack_through_the_public_class(engine):
    import contextlib, io
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    sc = make_scene('wavy', 5, 96, 72, seed=3, arc_deg=6.0)
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
    with contextlib.redirect_stdout(io.StringIO()):
        pm = PatchMatchMVS(cam, scale=1.0, patch_size=7, num_iterations=2, num_samples=2, min_views=2, seed=9,
                           schedule='redblack')
        p1, c1 = pm.reconstruct(sc.images, poses)
        p2, c2 = pm.reconstruct(sc.images, poses)
    assert p1.shape == p2.shape and np.array_equal(p1, p2) and np.array_equal(c1, c2)     # counter-based RNG: repeatable
    with pytest.raises(ValueError):
        PatchMatchMVS(cam, schedule='zigzag')
