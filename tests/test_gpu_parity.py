"""GPU (-m gpu): the CUDA path, called through the C ABI, against
  (1) the golden fixtures produced by the unmodified reference, and
  (2) the C oracle on the same seeded inputs (bit-exact: both use IEEE-RN fp32 in the same order).
Tolerances are the north_star's: NCC volumes 1e-4 absolute (observed <= 2.4e-7 vs the reference,
0 vs the oracle), argmax indices / masks / point counts bit-exact, PatchMatch >= 99.5 % of pixels
within 0.5 % depth and 1 degree normal (observed 100 %, depth bit-exact)."""
import numpy as np
import pytest
import torch

from helpers import engine_cams, load, oracle_cams, pm_agreement, replay_rng
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

SWEEPS = ['planesweep_plane_64x48_k5.npz', 'planesweep_wavy_80x60_k7.npz', 'planesweep_tt_161x97_k7.npz']
PMS = ['patchmatch_wavy_64x48_k11.npz', 'patchmatch_plane_57x43_k5.npz', 'patchmatch_tt_160x120_k11.npz']


def dev(engine, a, dtype=torch.float32):
    return engine.to_device(np.asarray(a), dtype)


@pytest.mark.parametrize('k', [3, 5, 7, 9, 11, 13])
@pytest.mark.parametrize('shape', [(48, 64), (97, 161), (33, 70)])
def test_ref_stats_bitexact(engine, k, shape):
    rng = np.random.default_rng(k)
    x = rng.random(shape, dtype=np.float32)
    mean1, var1 = engine.ref_stats(dev(engine, x), k)
    m = orc.box_mean(x, k)
    v = orc.box_mean(x * x, k) - m * m
    assert np.array_equal(mean1.cpu().numpy(), m)
    assert np.array_equal(var1.cpu().numpy(), v)


@pytest.mark.parametrize('name', SWEEPS)
def test_plane_sweep_vs_reference_and_oracle(engine, name):
    g = load(name)
    ref, nb = int(g['ref_idx']), list(g['neighbors'])
    k, th = int(g['patch']), float(g['thresh'])
    d32 = g['depths'].astype(np.float32)
    r = engine.plane_sweep(dev(engine, g['gray'][ref]), [dev(engine, g['gray'][i]) for i in nb],
                           engine_cams(g, ref, nb), d32, k, th, dump_ncc=True)
    depth_map, conf = r['depth_map'].cpu().numpy(), r['confidence'].cpu().numpy()
    idx, ncc = r['best_idx'].cpu().numpy(), r['ncc'].cpu().numpy()
    # vs reference fixture
    assert np.array_equal(depth_map, g['depth_map'])
    assert np.array_equal(conf, g['confidence'])
    if 'ncc' in g.files:
        assert np.abs(ncc - g['ncc']).max() <= 1e-4
    # vs oracle: everything bit-exact
    o = orc.plane_sweep(g['gray'][ref], [g['gray'][i] for i in nb], oracle_cams(orc, g, ref, nb), g['depths'], k, th,
                        dump_ncc=True)
    assert np.array_equal(ncc, o['ncc'])
    assert np.array_equal(idx, o['best_idx'])
    # production path (no NCC dump): the threshold test is decided without sqrt/div away from the threshold
    q = engine.plane_sweep(dev(engine, g['gray'][ref]), [dev(engine, g['gray'][i]) for i in nb],
                           engine_cams(g, ref, nb), d32, k, th)
    assert np.array_equal(q['best_idx'].cpu().numpy(), o['best_idx'])
    assert np.array_equal(q['confidence'].cpu().numpy(), g['confidence'])
    assert np.array_equal(q['depth_map'].cpu().numpy(), g['depth_map'])
    # fusion: mask, count, order, colours exact; coordinates to 1e-9
    pts, col = engine.mask_backproject(r['depth_map'], r['confidence'], dev(engine, g['bgr'][ref], torch.uint8),
                                       np.linalg.inv(g['K_scaled']), g['R'][ref], g['t'][ref],
                                       float(g['min_confidence']), 1)
    assert tuple(pts.shape) == g['points'].shape
    assert np.array_equal(col.cpu().numpy(), g['colors'])
    np.testing.assert_allclose(pts.cpu().numpy(), g['points'], rtol=0, atol=1e-9)


@pytest.mark.parametrize('name', PMS)
def test_patchmatch_vs_reference_and_oracle(engine, name):
    g = load(name)
    ref, src = int(g['ref_idx']), list(g['sources'])
    rng = {k: g[k] for k in ('depth0', 'normal0', 'ref_u', 'ref_n')} if 'ref_u' in g.files else replay_rng(g)
    if rng is None:
        pytest.skip('torch CPU generator on this host does not reproduce the recorded RNG stream')
    k, iters, samples = int(g['patch']), int(g['iters']), int(g['samples'])
    dmin, dmax = float(g['depth_min']), float(g['depth_max'])
    r = engine.pm_view(dev(engine, g['gray'][ref]), [dev(engine, g['gray'][i]) for i in src], engine_cams(g, ref, src),
                       k, iters, samples, dmin, dmax, init_depth=dev(engine, rng['depth0']),
                       init_normal=dev(engine, rng['normal0']), rand_u=dev(engine, rng['ref_u']),
                       rand_n=dev(engine, rng['ref_n']), dump_costs=True, want_best=True)
    depth, normal = r['depth'].cpu().numpy(), r['normal'].cpu().numpy()
    conf, costs = r['confidence'].cpu().numpy(), r['costs'].cpu().numpy()
    # vs reference fixture (north_star gate: >= 99.5 %)
    assert pm_agreement(depth, normal, g['depth'], g['normal']) >= 0.995
    assert np.array_equal(depth, g['depth'])
    assert np.array_equal(conf, g['confidence'])
    if 'costs' in g.files:
        fin = np.isfinite(g['costs'])
        assert np.array_equal(np.isfinite(costs), fin)
        assert np.abs(costs[fin] - g['costs'][fin]).max() <= 1e-4
    # vs oracle: bit-exact incl. every intermediate cost map (NaN-aware)
    o = orc.pm_view(g['gray'][ref], [g['gray'][i] for i in src], oracle_cams(orc, g, ref, src), k, iters, samples,
                    dmin, dmax, rng['depth0'], rng['normal0'], rng['ref_u'], rng['ref_n'], dump_costs=True)
    assert np.array_equal(costs, o['costs'], equal_nan=True)
    assert np.array_equal(depth, o['depth'])
    assert np.array_equal(normal, o['normal'])
    assert np.array_equal(conf, o['confidence'])
    assert np.array_equal(r['best_cost'].cpu().numpy(), o['best_cost'], equal_nan=True)
    # fusion
    pts, col = engine.mask_backproject(r['depth'], r['confidence'], dev(engine, g['bgr'][ref], torch.uint8),
                                       np.linalg.inv(g['K_scaled']), g['R'][ref], g['t'][ref], float(g['min_views']), 0)
    assert tuple(pts.shape) == g['points'].shape
    assert np.array_equal(col.cpu().numpy(), g['colors'])
    np.testing.assert_allclose(pts.cpu().numpy(), g['points'], rtol=0, atol=1e-9)


@pytest.mark.parametrize('k', [3, 5, 7, 9, 11, 13])
def test_pm_cost_and_confidence_all_patch_sizes(engine, k):
    """Seeded inputs at a ragged size (not a multiple of the 64x32 tile), every supported window."""
    from b200dense.scenes import gray_f32, make_scene
    from b200dense.engine import kinv32, make_cams
    sc = make_scene('wavy', 5, 150, 83, seed=k)
    gray = gray_f32(sc)
    ref, src = 2, [1, 3, 0, 4]
    Ki = kinv32(sc.K)
    R = [sc.poses[i][0] for i in range(5)]
    t = [sc.poses[i][1] for i in range(5)]
    rng = np.random.default_rng(k)
    depth = rng.uniform(sc.depth_range[0], sc.depth_range[1], size=gray[0].shape).astype(np.float32)
    depth[:, :40] = sc.depth_gt[ref][:, :40]                         # part of the map at the true depth
    ocams = orc.Cams(sc.K, Ki, R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    ecams = make_cams(sc.K, Ki, R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    for mode in (0, 1):
        got = engine.pm_cost(dev(engine, gray[ref]), [dev(engine, gray[i]) for i in src], ecams, dev(engine, depth), k, mode)
        exp = orc.pm_cost(gray[ref], [gray[i] for i in src], ocams, depth, k, mode)
        assert np.array_equal(got.cpu().numpy(), exp, equal_nan=True), f'mode {mode}'


def test_degenerate_inputs(engine):
    """Flat images (zero variance -> NaN cost that never wins), far depths (< 2 valid sources -> +inf)."""
    g = load('patchmatch_plane_57x43_k5.npz')
    ref, src = int(g['ref_idx']), list(g['sources'])
    H, W = g['gray'][0].shape
    ecams, ocams = engine_cams(g, ref, src), oracle_cams(orc, g, ref, src)
    far = np.full((H, W), 1e6, np.float32)
    a = engine.pm_cost(dev(engine, g['gray'][ref]), [dev(engine, g['gray'][i]) for i in src], ecams, dev(engine, far), 5)
    b = orc.pm_cost(g['gray'][ref], [g['gray'][i] for i in src], ocams, far, 5)
    assert np.array_equal(a.cpu().numpy(), b, equal_nan=True) and np.isinf(b).any()
    z = np.zeros((H, W), np.float32)
    d5 = np.full((H, W), 5.0, np.float32)
    a = engine.pm_cost(dev(engine, z), [dev(engine, z) for _ in src], ecams, dev(engine, d5), 5)
    b = orc.pm_cost(z, [z for _ in src], ocams, d5, 5)
    assert np.array_equal(a.cpu().numpy(), b, equal_nan=True)


def test_unsupported_arguments_raise(engine):
    from b200dense import B200DenseError
    x = torch.zeros((16, 16), device=engine.device)
    with pytest.raises(B200DenseError):
        engine.ref_stats(x, 4)
    with pytest.raises(B200DenseError):
        engine.ref_stats(x, 15)


def test_exact_division_fallback_paths(engine):
    """Operands outside the range where the shared-reciprocal division is trusted (zero numerators at the
    principal point, z + 1e-8 near zero, huge / tiny / negative / inf / NaN depths) must take the exact
    twin and still match the oracle bit for bit."""
    from b200dense.engine import make_cams
    H, W, k = 48, 64, 5
    rng = np.random.default_rng(3)
    K = np.array([[50.0, 0, 32.0], [0, 50.0, 24.0], [0, 0, 1.0]])          # integer principal point -> exact zeros
    Ki = np.linalg.inv(K).astype(np.float32)
    I3 = np.eye(3)
    R = [I3, I3, I3, I3]
    t = [np.zeros(3), np.array([0.0, 0.0, -2.0]), np.array([0.0, 0.0, 1.0]), np.array([0.25, 0.0, 0.0])]
    gray = [rng.random((H, W), dtype=np.float32) for _ in range(4)]
    depth = rng.uniform(1.0, 4.0, size=(H, W)).astype(np.float32)
    depth[0, :8] = [0.0, 1e-30, 1e30, -3.0, np.inf, np.nan, 2.0, 1e-12]
    depth[10:20, :] = 2.0                                                    # z = 2 - 2 = 0 in source 1: zz = 1e-8
    depth[24, 32] = 1.0                                                      # principal point: numerators exactly 0
    src = [1, 2, 3]
    ocams = orc.Cams(K, Ki, R[0], t[0], [R[i] for i in src], [t[i] for i in src])
    ecams = make_cams(K, Ki, R[0], t[0], [R[i] for i in src], [t[i] for i in src])
    for mode in (0, 1):
        got = engine.pm_cost(dev(engine, gray[0]), [dev(engine, gray[i]) for i in src], ecams, dev(engine, depth), k, mode)
        exp = orc.pm_cost(gray[0], [gray[i] for i in src], ocams, depth, k, mode)
        assert np.array_equal(got.cpu().numpy(), exp, equal_nan=True), f'mode {mode}'


@pytest.mark.parametrize('n_src', [1, 2, 3, 6, 7, 8])
def test_every_source_count_bitexact(engine, n_src):
    """sample_kernel has compile-time-unrolled variants for 4/5/6 sources and a run-time loop for the rest; the NCC
    kernel double-buffers its TMA stages over any count up to B200D_MAX_SRC.  Cost, confidence, a PatchMatch view and a
    plane sweep against the oracle for the counts the fixtures do not cover."""
    from b200dense.engine import kinv32, make_cams
    from b200dense.scenes import gray_f32, make_scene
    sc = make_scene('wavy', 9, 96, 72, seed=11 + n_src, arc_deg=4.0)
    gray = gray_f32(sc)
    ref, src = 4, [3, 5, 2, 6, 1, 7, 0, 8][:n_src]
    Ki = kinv32(sc.K)
    R = [sc.poses[i][0] for i in range(9)]
    t = [sc.poses[i][1] for i in range(9)]
    ecams = make_cams(sc.K, Ki, R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    ocams = orc.Cams(sc.K, Ki, R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    H, W = gray[0].shape
    dmin, dmax = sc.depth_range
    g = np.random.default_rng(n_src)
    depth = (g.random((H, W), dtype=np.float32) * np.float32(dmax - dmin) + np.float32(dmin))
    d_ref, d_src = dev(engine, gray[ref]), [dev(engine, gray[i]) for i in src]
    for mode in (0, 1):
        a = engine.pm_cost(d_ref, d_src, ecams, dev(engine, depth), 7, mode).cpu().numpy()
        b = orc.pm_cost(gray[ref], [gray[i] for i in src], ocams, depth, 7, mode)
        assert np.array_equal(a, b, equal_nan=True)
    iters, samples = 2, 2
    n0 = g.standard_normal((H, W, 3), dtype=np.float32)
    n0 /= np.linalg.norm(n0, axis=-1, keepdims=True)
    ru = g.random((iters * samples, H, W), dtype=np.float32)
    rn = g.standard_normal((iters * samples, H, W, 3), dtype=np.float32)
    out = engine.pm_view(d_ref, d_src, ecams, 7, iters, samples, dmin, dmax, init_depth=dev(engine, depth),
                         init_normal=dev(engine, n0), rand_u=dev(engine, ru), rand_n=dev(engine, rn))
    exp = orc.pm_view(gray[ref], [gray[i] for i in src], ocams, 7, iters, samples, dmin, dmax, depth, n0, ru, rn)
    for key in ('depth', 'normal', 'confidence'):
        assert np.array_equal(out[key].cpu().numpy(), exp[key], equal_nan=True), key
    depths = (1.0 / np.linspace(1 / dmax, 1 / dmin, 12)).astype(np.float32)
    sw = engine.plane_sweep(d_ref, d_src, ecams, depths, 5, 0.8)
    ex = orc.plane_sweep(gray[ref], [gray[i] for i in src], ocams, depths, 5, 0.8)
    assert np.array_equal(sw['best_idx'].cpu().numpy(), ex['best_idx'])
    assert np.array_equal(sw['confidence'].cpu().numpy(), ex['confidence'])
