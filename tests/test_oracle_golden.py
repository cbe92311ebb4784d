"""CPU: the C oracle against fixtures produced by EXECUTING the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle.

Bit-exact: window sums, reprojection, sampling, inf/NaN pattern, depth, confidence, argmax index,
masks, point count/order, colours.  NCC / cost values: <= 2.4e-7 absolute -- the reference's
torch.sqrt goes through MKL VML (vsSqrt, HA mode), which is 1 ulp low in ~0.6 % of calls; the
oracle (and the CUDA kernels) use the correctly rounded IEEE sqrt.  north_star tolerance: 1e-4.
"""
import numpy as np
import pytest

from helpers import load, oracle_cams, pm_agreement, replay_rng
from oracle import oracle as orc

NCC_TOL = 5e-7          # observed max 2.4e-7; required by north_star: 1e-4

SWEEPS = ['planesweep_plane_64x48_k5.npz', 'planesweep_wavy_80x60_k7.npz', 'planesweep_tt_161x97_k7.npz',
          'planesweep_plane_60x44_k9.npz', 'planesweep_wavy_72x54_k13.npz']
PMS = ['patchmatch_wavy_64x48_k11.npz', 'patchmatch_plane_57x43_k5.npz', 'patchmatch_tt_160x120_k11.npz',
       'patchmatch_wavy_60x45_k9.npz', 'patchmatch_plane_64x48_k13.npz']


@pytest.mark.parametrize('k', [3, 5, 7, 9, 11, 13])
def test_box_mean_matches_conv2d_bitexact(k):
    """F.conv2d(x, ones/k^2, padding=k//2) == row-major FMA chain (dense_stereo.py:325-341)."""
    import torch
    import torch.nn.functional as F
    torch.manual_seed(k)
    x = torch.rand(97, 131)
    ref = F.conv2d(x[None, None], torch.ones(1, 1, k, k) / (k * k), padding=k // 2)[0, 0].numpy()
    got = orc.box_mean(x.numpy(), k)
    if k == 3:
        # For 3x3 on small images PyTorch picks a different conv2d kernel (summation order differs by
        # <= 1 ulp; at 480x640 it is bit-exact again).  k=3 is not used by any reference default.
        assert np.abs(got - ref).max() <= 1.2e-7
    else:
        assert np.array_equal(got, ref)


@pytest.mark.parametrize('name', SWEEPS)
def test_plane_sweep_oracle_vs_reference(name):
    g = load(name)
    ref, nb = int(g['ref_idx']), list(g['neighbors'])
    r = orc.plane_sweep(g['gray'][ref], [g['gray'][i] for i in nb], oracle_cams(orc, g, ref, nb), g['depths'],
                        int(g['patch']), float(g['thresh']), dump_ncc='ncc' in g.files)
    assert np.array_equal(r['depth_map'], g['depth_map'])          # argmax depth: bit-exact
    assert np.array_equal(r['confidence'], g['confidence'])
    assert np.array_equal(g['depths'].astype(np.float32)[r['best_idx']], g['depth_map'])
    if 'ncc' in g.files:
        assert np.abs(r['ncc'] - g['ncc']).max() <= NCC_TOL
        assert (r['ncc'] != g['ncc']).mean() < 0.02
    pts, col = orc.mask_backproject(r['depth_map'], r['confidence'], g['bgr'][ref], np.linalg.inv(g['K_scaled']),
                                    g['R'][ref], g['t'][ref], float(g['min_confidence']), 1)
    assert pts.shape == g['points'].shape                           # mask / count: bit-exact
    assert np.array_equal(col, g['colors'])
    np.testing.assert_allclose(pts, g['points'], rtol=0, atol=1e-9)


@pytest.mark.parametrize('name', PMS)
def test_patchmatch_oracle_vs_reference(name):
    g = load(name)
    ref, src = int(g['ref_idx']), list(g['sources'])
    rng = {k: g[k] for k in ('depth0', 'normal0', 'ref_u', 'ref_n')} if 'ref_u' in g.files else replay_rng(g)
    if rng is None:
        pytest.skip('torch CPU generator on this host does not reproduce the recorded RNG stream')
    dump = 'costs' in g.files
    r = orc.pm_view(g['gray'][ref], [g['gray'][i] for i in src], oracle_cams(orc, g, ref, src), int(g['patch']),
                    int(g['iters']), int(g['samples']), float(g['depth_min']), float(g['depth_max']),
                    rng['depth0'], rng['normal0'], rng['ref_u'], rng['ref_n'], dump_costs=dump)
    if dump:
        a, b = r['costs'], g['costs']
        assert np.array_equal(np.isinf(a), np.isinf(b)) and np.array_equal(np.isnan(a), np.isnan(b))
        fin = np.isfinite(b)
        assert np.abs(a[fin] - b[fin]).max() <= NCC_TOL
    assert np.array_equal(r['depth'], g['depth'])                    # selected hypotheses: bit-exact
    assert np.array_equal(r['confidence'], g['confidence'])
    assert np.abs(r['normal'] - g['normal']).max() < 2e-6
    assert pm_agreement(r['depth'], r['normal'], g['depth'], g['normal']) == 1.0
    pts, col = orc.mask_backproject(r['depth'], r['confidence'], g['bgr'][ref], np.linalg.inv(g['K_scaled']),
                                    g['R'][ref], g['t'][ref], float(g['min_views']), 0)
    assert pts.shape == g['points'].shape and np.array_equal(col, g['colors'])
    np.testing.assert_allclose(pts, g['points'], rtol=0, atol=1e-9)


def test_stored_rng_matches_replay():
    """The fixture that stores its RNG tensors must agree with the seed replay used by the larger one."""
    g = load('patchmatch_wavy_64x48_k11.npz')
    rng = replay_rng(g)
    if rng is None:
        pytest.skip('RNG stream not reproducible on this host')
    for k in ('depth0', 'normal0', 'ref_u', 'ref_n'):
        assert np.array_equal(rng[k], g[k]), k


def test_cost_edge_cases():
    """Ragged / degenerate inputs: fewer than two valid sources -> +inf; zero-variance patches ->
    NaN cost that never wins the strict `<` (mvs_patchmatch.py:388, 410, 452)."""
    g = load('patchmatch_plane_57x43_k5.npz')
    ref, src = int(g['ref_idx']), list(g['sources'])
    H, W = g['gray'][0].shape
    cams = oracle_cams(orc, g, ref, src)
    far = np.full((H, W), 1e6, np.float32)          # projects (almost) at the epipole: mostly invalid
    c = orc.pm_cost(g['gray'][ref], [g['gray'][i] for i in src], cams, far, 5, 0)
    assert np.isinf(c).any()
    flat = [np.zeros((H, W), np.float32) for _ in src]
    c2 = orc.pm_cost(np.zeros((H, W), np.float32), flat, cams, np.full((H, W), 5.0, np.float32), 5, 0)
    assert (np.isnan(c2) | np.isinf(c2) | (c2 == 1.0)).all()
    conf = orc.pm_cost(g['gray'][ref], [g['gray'][i] for i in src], cams, g['depth'], 5, 1)
    assert np.array_equal(conf, g['confidence'])
