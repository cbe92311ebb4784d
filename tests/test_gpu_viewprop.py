"""GPU (-m gpu): the opt-in view-propagation schedule (EXTENSION -- the reference's README names the step, README.md:117,
its code has none; no dackey-wav oracle).  The definition is restated here in NumPy on top of the oracle's
cost evaluation (`_compute_patch_cost` arithmetic, pinned) and of oracle.viewprop_candidates (the proposal rule, written
with the pinned primitives); the kernel path -- staged b200d_pm_view + b200d_pm_viewprop, maps exchanged through
all_gather_maps -- must reproduce it bit for bit, and the reference schedule must be untouched."""
import contextlib
import io

import numpy as np
import pytest
import torch

from b200dense import _cabi
from b200dense.engine import kinv32, make_cams
from b200dense.scenes import gray_f32, make_scene
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
F = np.float32


def _shift(a, dy, dx, fill):
    H, W = a.shape[:2]
    out = np.full_like(a, fill)
    ys, yd = (slice(dy, H), slice(0, H - dy)) if dy >= 0 else (slice(0, H + dy), slice(-dy, H))
    xs, xd = (slice(dx, W), slice(0, W - dx)) if dx >= 0 else (slice(0, W + dx), slice(-dx, W))
    out[yd, xd] = a[ys, xs]
    return out


class NumpyView:
    """One view's state advanced with NumPy selection on top of the oracle's cost evaluation."""

    def __init__(self, gray, ref, src, ocams, k, dmin, dmax, depth0, normal0, ru, rn, samples):
        self.gray, self.ref, self.src, self.ocams, self.k = gray, ref, src, ocams, k
        self.dmin, self.dmax, self.ru, self.rn, self.samples = dmin, dmax, ru, rn, samples
        self.depth, self.normal = depth0.copy(), normal0.copy()
        self.best = np.full(depth0.shape, np.inf, F)

    def cost(self, d, mode=0):
        return orc.pm_cost(self.gray[self.ref], [self.gray[i] for i in self.src], self.ocams, np.ascontiguousarray(d),
                           self.k, mode)

    def adopt(self, win, cd, cn, cost):
        self.depth[win], self.normal[win], self.best[win] = cd[win], cn[win], cost[win]

    def iteration(self, it):
        """The reference's iteration (mvs_patchmatch.py:415-491, slicing semantics of SURVEY F3)."""
        sgn = 1 if it % 2 == 0 else -1
        for dy, dx in ((sgn, 0), (0, sgn)):
            cd, cn = _shift(self.depth, dy, dx, F(self.dmin)), _shift(self.normal, dy, dx, F(0))
            c = self.cost(cd)
            self.adopt(c < self.best, cd, cn, c)
        dr, nr = F((self.dmax - self.dmin) * 0.5 ** it), F(0.5 * 0.5 ** it)
        for sm in range(self.samples):
            u, r = self.ru[it * self.samples + sm], self.rn[it * self.samples + sm]
            cd = np.clip(self.depth + (u * F(2) - F(1)) * dr, F(self.dmin), F(self.dmax)).astype(F)
            c3 = self.normal + r * nr
            nn = np.sqrt((c3[..., 0] * c3[..., 0] + c3[..., 1] * c3[..., 1]) + c3[..., 2] * c3[..., 2])
            cn = c3 / np.maximum(nn, F(1e-12))[..., None]
            c = self.cost(cd)
            self.adopt(c < self.best, cd, cn, c)

    def viewprop(self, snap_depth, snap_normal):
        """One proposal per source, in source order, from the SNAPSHOT of the sources' maps."""
        costs = []
        for s, v in enumerate(self.src):
            cd, cn, allow = orc.viewprop_candidates(self.depth, self.normal, snap_depth[v], snap_normal[v], s, self.ocams,
                                                    self.dmin, self.dmax)
            c = self.cost(cd)
            costs.append(c)
            self.adopt((c < self.best) & allow, cd, cn, c)
        return costs


def _setup(W, H, n_views, seed):
    sc = make_scene('wavy', n_views, W, H, seed=seed, arc_deg=5.0)
    gray = gray_f32(sc)
    Ki = kinv32(sc.K)
    R = [sc.poses[i][0] for i in range(n_views)]
    t = [sc.poses[i][1] for i in range(n_views)]
    srcs = {r: [i for i in sorted(range(n_views), key=lambda j: (abs(j - r), j)) if i != r][:3] for r in range(n_views)}
    cams = {r: (sc.K, Ki, R[r], t[r], [R[i] for i in srcs[r]], [t[i] for i in srcs[r]]) for r in range(n_views)}
    return sc, gray, srcs, cams


@pytest.mark.parametrize('shape,k', [((64, 48), 7), ((57, 43), 5)])
def test_viewprop_schedule_equals_its_definition(engine, shape, k):
    W, H = shape
    n_views, iters, samples = 4, 2, 2
    sc, gray, srcs, cams = _setup(W, H, n_views, seed=17)
    dmin, dmax = sc.depth_range
    g = np.random.default_rng(9)
    init = {}
    for r in range(n_views):
        d0 = (g.random((H, W), dtype=F) * F(dmax - dmin) + F(dmin))
        n0 = g.standard_normal((H, W, 3), dtype=F)
        n0 /= np.linalg.norm(n0, axis=-1, keepdims=True)
        init[r] = (d0, n0, g.random((iters * samples, H, W), dtype=F), g.standard_normal((iters * samples, H, W, 3), dtype=F))
    # ---- definition
    views = {r: NumpyView(gray, r, srcs[r], orc.Cams(*cams[r]), k, dmin, dmax, *init[r], samples) for r in range(n_views)}
    vp_costs = {}
    for it in range(iters):
        for r in views:
            views[r].iteration(it)
        snap_d = {r: views[r].depth.copy() for r in views}
        snap_n = {r: views[r].normal.copy() for r in views}
        for r in views:
            vp_costs[(it, r)] = views[r].viewprop(snap_d, snap_n)
    # ---- kernel path: staged PatchMatch + proposals from snapshots
    d = engine.to_device
    dg = {i: d(gray[i]) for i in range(n_views)}
    st = {r: dict(depth=torch.empty((H, W), device=engine.device), normal=torch.empty((H, W, 3), device=engine.device),
                  confidence=torch.empty((H, W), device=engine.device), best_cost=torch.empty((H, W), device=engine.device))
          for r in range(n_views)}
    kw = {r: dict(init_depth=d(init[r][0]), init_normal=d(init[r][1]), rand_u=d(init[r][2]), rand_n=d(init[r][3]))
          for r in range(n_views)}
    adopted = 0
    for it in range(iters):
        for r in range(n_views):
            stage = _cabi.STAGE_NO_CONF | (_cabi.STAGE_NO_INIT if it else 0)
            engine.pm_view(dg[r], [dg[i] for i in srcs[r]], make_cams(*cams[r]), k, iters, samples, dmin, dmax, out=st[r],
                           it_begin=it, it_end=it + 1, stage=stage, **kw[r])
        snap_d = {r: st[r]['depth'].clone() for r in st}
        snap_n = {r: st[r]['normal'].clone() for r in st}
        for r in range(n_views):
            before = st[r]['depth'].clone()
            costs = engine.pm_viewprop(dg[r], [dg[i] for i in srcs[r]], make_cams(*cams[r]), k, dmin, dmax,
                                       [snap_d[i] for i in srcs[r]], [snap_n[i] for i in srcs[r]], st[r], dump_costs=True)
            for s in range(len(srcs[r])):
                assert np.array_equal(costs[s].cpu().numpy(), vp_costs[(it, r)][s], equal_nan=True), (it, r, s)
            adopted += int((before != st[r]['depth']).sum().item())
    for r in range(n_views):
        engine.pm_view(dg[r], [dg[i] for i in srcs[r]], make_cams(*cams[r]), k, iters, samples, dmin, dmax, out=st[r],
                       it_begin=iters, it_end=0, stage=_cabi.STAGE_NO_INIT, **kw[r])
        assert np.array_equal(st[r]['depth'].cpu().numpy(), views[r].depth), r
        assert np.array_equal(st[r]['normal'].cpu().numpy(), views[r].normal), r
        assert np.array_equal(st[r]['best_cost'].cpu().numpy(), views[r].best, equal_nan=True), r
        assert np.array_equal(st[r]['confidence'].cpu().numpy(), views[r].cost(views[r].depth, 1)), r
    assert adopted > 0                                        # proposals were really adopted somewhere
    # ---- staged execution without proposals == the monolithic, graded schedule (bit-exact vs the oracle)
    r = 1
    o = orc.pm_view(gray[r], [gray[i] for i in srcs[r]], orc.Cams(*cams[r]), k, iters, samples, dmin, dmax, *init[r])
    st2 = {key: torch.empty_like(v) for key, v in st[r].items()}
    for it in range(iters):
        engine.pm_view(dg[r], [dg[i] for i in srcs[r]], make_cams(*cams[r]), k, iters, samples, dmin, dmax, out=st2,
                       it_begin=it, it_end=it + 1, stage=_cabi.STAGE_NO_CONF | (_cabi.STAGE_NO_INIT if it else 0), **kw[r])
    engine.pm_view(dg[r], [dg[i] for i in srcs[r]], make_cams(*cams[r]), k, iters, samples, dmin, dmax, out=st2,
                   it_begin=iters, it_end=0, stage=_cabi.STAGE_NO_INIT, **kw[r])
    for key, ok in (('depth', 'depth'), ('normal', 'normal'), ('confidence', 'confidence'), ('best_cost', 'best_cost')):
        assert np.array_equal(st2[key].cpu().numpy(), o[ok], equal_nan=True), key


def test_viewprop_through_the_public_class(engine):
    """schedule='viewprop' runs end to end, is repeatable (counter-based RNG), differs from the reference schedule, and
    with ground-truth-like convergence proposes real hypotheses (clouds are non-empty)."""
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    sc = make_scene('wavy', 6, 96, 72, seed=3, arc_deg=6.0)
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
    sparse = np.stack([np.zeros(20), np.zeros(20), np.linspace(4.0, 6.0, 20)], axis=1)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for sched in ('reference', 'viewprop'):
            pm = PatchMatchMVS(cam, scale=1.0, patch_size=7, num_iterations=2, num_samples=2, min_views=2, seed=9,
                               schedule=sched)
            out[sched] = [pm.reconstruct(sc.images, poses, sparse_points=sparse) for _ in range(2)]
    for sched in out:
        (p1, c1), (p2, c2) = out[sched]
        assert len(p1) > 0 and np.array_equal(p1, p2) and np.array_equal(c1, c2)
    assert not (out['reference'][0][0].shape == out['viewprop'][0][0].shape
                and np.array_equal(out['reference'][0][0], out['viewprop'][0][0]))
