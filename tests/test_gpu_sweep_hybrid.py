"""GPU (-m gpu): the plane sweep's bounded decision (csrc/fast_kernels.cuh, ARITH = A_HYBRID) must be invisible.

In B200D_PARITY mode `ncc > thresh` (dense_stereo.py:303) is decided from separable window sums plus a rigorous error
bound; undecided outputs redo the reference's chains.  Counts, arg-max and depth must therefore equal, bit for bit,
(1) the C oracle and (2) B200D_PARITY_CHAINS (the chains evaluated for every output) -- for every window size,
neighbour count, threshold, on ragged sizes, flat / dark / saturated / non-finite images, and at the benchmark shape."""
import numpy as np
import pytest
import torch

from b200dense import PARITY, PARITY_CHAINS, PARITY_CHECK
from b200dense.engine import kinv32, make_cams
from b200dense.scenes import gray_f32, make_scene
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
F = np.float32


def _scene(W, H, n=7, seed=3, kind='wavy'):
    sc = make_scene(kind, n, W, H, seed=seed, arc_deg=5.0)
    return sc, gray_f32(sc)


def _cams(sc, ref, nb):
    Ki = kinv32(sc.K)
    R = {i: sc.poses[i][0] for i in sc.poses}
    t = {i: sc.poses[i][1] for i in sc.poses}
    args = (sc.K, Ki, R[ref], t[ref], [R[i] for i in nb], [t[i] for i in nb])
    return make_cams(*args), orc.Cams(*args)


def _sweep_all(engine, gray, ref, nb, ecams, depths, k, th):
    d = engine.to_device
    a = (d(gray[ref]), [d(gray[i]) for i in nb], ecams, depths, k, th)
    engine.sweep_fallbacks(reset=True)
    hy = engine.plane_sweep(*a, arith=PARITY)
    n_fb = engine.sweep_fallbacks()
    ch = engine.plane_sweep(*a, arith=PARITY_CHAINS)
    return hy, ch, n_fb


def _same(a, b):
    for key in ('best_idx', 'confidence', 'depth_map'):
        x = a[key].cpu().numpy() if torch.is_tensor(a[key]) else a[key]
        y = b[key].cpu().numpy() if torch.is_tensor(b[key]) else b[key]
        assert np.array_equal(x, y), key


@pytest.mark.parametrize('k', [3, 5, 7, 9, 11, 13])
@pytest.mark.parametrize('shape,n_nb', [((150, 83), 4), ((64, 32), 1), ((131, 70), 6)])
def test_hybrid_equals_chains_and_oracle(engine, k, shape, n_nb):
    W, H = shape
    sc, gray = _scene(W, H, seed=k)
    ref, nb = 3, [2, 4, 1, 5, 0, 6][:n_nb]
    ecams, ocams = _cams(sc, ref, nb)
    dmin, dmax = sc.depth_range
    depths = (1.0 / np.linspace(1 / dmax, 1 / dmin, 24)).astype(F)
    for th in (0.8, 0.5, 0.97):
        hy, ch, _ = _sweep_all(engine, gray, ref, nb, ecams, depths, k, th)
        _same(hy, ch)
        o = orc.plane_sweep(gray[ref], [gray[i] for i in nb], ocams, depths, k, th)
        _same(hy, o)


def test_hybrid_eight_neighbours_and_disabled_quick_test(engine):
    sc, gray = _scene(120, 90, n=9, seed=11)
    ref, nb = 4, [3, 5, 2, 6, 1, 7, 0, 8]
    ecams, ocams = _cams(sc, ref, nb)
    depths = (1.0 / np.linspace(1 / sc.depth_range[1], 1 / sc.depth_range[0], 16)).astype(F)
    hy, ch, _ = _sweep_all(engine, gray, ref, nb, ecams, depths, 7, 0.8)
    _same(hy, ch)
    _same(hy, orc.plane_sweep(gray[ref], [gray[i] for i in nb], ocams, depths, 7, 0.8))
    # thresholds outside (1e-3, 1e3) switch the quick tests off: every comparison is the exact sqrt + division
    for th in (1e-4, -0.5, 0.0):
        hy, ch, _ = _sweep_all(engine, gray, ref, nb[:4], make_cams(*_args(sc, ref, nb[:4])), depths, 5, th)
        _same(hy, ch)
        _same(hy, orc.plane_sweep(gray[ref], [gray[i] for i in nb[:4]], orc.Cams(*_args(sc, ref, nb[:4])), depths, 5, th))


def _args(sc, ref, nb):
    Ki = kinv32(sc.K)
    return (sc.K, Ki, sc.poses[ref][0], sc.poses[ref][1], [sc.poses[i][0] for i in nb], [sc.poses[i][1] for i in nb])


@pytest.mark.parametrize('case', ['flat', 'dark', 'steps', 'saturated', 'tiny', 'nan', 'inf', 'negative', 'huge'])
def test_hybrid_on_degenerate_images(engine, case):
    """Where var = E[x^2] - E[x]^2 cancels (flat patches) the bound must send the output to the exact path, and
    non-finite samples must never be 'decided'."""
    sc, gray = _scene(96, 72, seed=7)
    g = {i: gray[i].copy() for i in range(len(gray))}
    rng = np.random.default_rng(1)
    for i in g:
        if case == 'flat':
            g[i][:, 30:70] = F(0.5)
        elif case == 'dark':
            g[i] *= F(1e-3)
        elif case == 'steps':
            g[i] = (np.floor(g[i] * 4) / 4).astype(F)
        elif case == 'saturated':
            g[i] = np.minimum(g[i] * 3, 1).astype(F)
        elif case == 'tiny':
            g[i] = (g[i] * F(1e-18)).astype(F)
        elif case == 'nan':
            g[i][rng.integers(0, 72, 20), rng.integers(0, 96, 20)] = np.nan
        elif case == 'inf':
            g[i][rng.integers(0, 72, 20), rng.integers(0, 96, 20)] = np.inf
        elif case == 'negative':
            g[i] = (g[i] - F(0.5)).astype(F)
        elif case == 'huge':
            g[i] = (g[i] * F(1e4)).astype(F)
    ref, nb = 3, [2, 4, 1, 5]
    ecams, ocams = _cams(sc, ref, nb)
    depths = (1.0 / np.linspace(1 / sc.depth_range[1], 1 / sc.depth_range[0], 20)).astype(F)
    for k in (5, 11):
        hy, ch, _ = _sweep_all(engine, g, ref, nb, ecams, depths, k, 0.8)
        _same(hy, ch)
        _same(hy, orc.plane_sweep(g[ref], [g[i] for i in nb], ocams, depths, k, 0.8))


def test_hybrid_at_benchmark_shape_reports_fallback_rate(engine):
    """1920x1080, N=6, k=7 on the benchmark's turntable (a slice of BASELINE configs[2]): hybrid == chains bit for bit;
    the share of outputs redone exactly is printed (expected: a fraction of a per cent)."""
    W, H, D, k = 1920, 1080, 256, 7                        # all 256 planes of BASELINE configs[2] for one reference view
    sc = make_scene('turntable', 64, W, H, arc_deg=360.0 / 64, seed=0, device='cuda', render=list(range(0, 10)))
    gray = gray_f32(sc)
    ref, nb = 3, [2, 4, 1, 5, 0, 6]
    ecams = make_cams(*_args(sc, ref, nb))
    depths = (1.0 / np.linspace(1 / sc.depth_range[1], 1 / sc.depth_range[0], D)).astype(F)
    hy, ch, n_fb = _sweep_all(engine, gray, ref, nb, ecams, depths, k, 0.8)
    _same(hy, ch)
    ref2, nb2 = 6, [5, 7, 4, 8, 3, 9]                     # a second view, every fourth plane
    hy2, ch2, _ = _sweep_all(engine, gray, ref2, nb2, make_cams(*_args(sc, ref2, nb2)), depths[::4].copy(), k, 0.8)
    _same(hy2, ch2)
    frac = n_fb / (W * H * len(nb) * D)
    print(f'hybrid sweep {W}x{H} D={D} N={len(nb)} k={k}: {n_fb} exact redos = {100 * frac:.3f} % of pixel x neighbour x plane outputs')
    assert 0 < frac < 0.05
    conf = hy['confidence'].cpu().numpy()
    assert conf.max() >= 3                                 # the sweep found consistent planes on this scene


@pytest.mark.parametrize('k', [3, 5, 7, 9, 11, 13])
def test_error_bound_is_never_exceeded(engine, k):
    """B200D_PARITY_CHECK evaluates, for EVERY output and neighbour, the reference's chains beside the separable sums and
    reports max |difference| / bound for box(s), box(s*s), box(r*s).  A ratio above 1 would falsify the derivation the
    bit-exact decision rests on (csrc/fast_kernels.cuh, sweep_decide_bounded).  Textured, flat, dark, denormal-small,
    huge and negative images; results must also equal the chains-everywhere twin."""
    sc, gray = _scene(150, 83, seed=k)
    ref, nb = 3, [2, 4, 1, 5, 0, 6]
    ecams, _ = _cams(sc, ref, nb)
    depths = (1.0 / np.linspace(1 / sc.depth_range[1], 1 / sc.depth_range[0], 16)).astype(F)
    d = engine.to_device
    worst = [0.0, 0.0, 0.0]
    for case in ('textured', 'flat', 'dark', 'tiny', 'huge', 'negative', 'steps'):
        g = {i: gray[i].copy() for i in range(len(gray))}
        for i in g:
            if case == 'flat':
                g[i][:, 40:100] = F(0.37)
            elif case == 'dark':
                g[i] *= F(3e-4)
            elif case == 'tiny':
                g[i] = (g[i] * F(1e-17)).astype(F)
            elif case == 'huge':
                g[i] = (g[i] * F(3e3)).astype(F)
            elif case == 'negative':
                g[i] = (g[i] - F(0.5)).astype(F)
            elif case == 'steps':
                g[i] = (np.floor(g[i] * 3) / 3).astype(F)
        a = (d(g[ref]), [d(g[i]) for i in nb], ecams, depths, k, 0.8)
        engine.sweep_bound_check(reset=True)
        chk = engine.plane_sweep(*a, arith=PARITY_CHECK)
        ratios = engine.sweep_bound_check()
        _same(chk, engine.plane_sweep(*a, arith=PARITY_CHAINS))
        assert all(np.isfinite(r) and 0.0 <= r <= 1.0 for r in ratios), (case, ratios)
        worst = [max(w, r) for w, r in zip(worst, ratios)]
    print(f'k={k}: max |separable - chain| / bound = m2 {worst[0]:.3f}, box(s*s) {worst[1]:.3f}, box(r*s) {worst[2]:.3f}')
    assert max(worst) > 0.0                                 # the check really compared something


def test_error_bound_at_benchmark_shape(engine):
    W, H, k = 1920, 1080, 7
    sc = make_scene('turntable', 64, W, H, arc_deg=360.0 / 64, seed=0, device='cuda', render=list(range(0, 8)))
    gray = gray_f32(sc)
    ref, nb = 3, [2, 4, 1, 5, 0, 6]
    depths = (1.0 / np.linspace(1 / sc.depth_range[1], 1 / sc.depth_range[0], 256)).astype(F)[::8].copy()
    d = engine.to_device
    engine.sweep_bound_check(reset=True)
    engine.plane_sweep(d(gray[ref]), [d(gray[i]) for i in nb], make_cams(*_args(sc, ref, nb)), depths, k, 0.8, arith=PARITY_CHECK)
    ratios = engine.sweep_bound_check()
    print(f'1920x1080, 32 planes x 6 neighbours = 398 M outputs: max |separable - chain| / bound = {ratios}')
    assert all(0.0 < r <= 1.0 for r in ratios), ratios
