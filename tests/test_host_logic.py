"""CPU: host-side logic of the drop-in classes against values produced by the reference
(tests/golden/host_heuristics.npz, oracle/make_golden.py): source-view / neighbour selection incl.
stable-sort tie behaviour, depth-range estimators, post filters, PLY format, CLI flags."""
import numpy as np
import pytest

from helpers import load
from core.camera import Camera, CameraPose, load_calibration
from core.dense_stereo import DenseStereoReconstructor
from core.mvs_patchmatch import DepthNormalMap, PatchMatchMVS
from core.utils import save_cameras_ply, save_ply


def _poses(g, tag):
    return {i: CameraPose(R=g[f'{tag}_R'][i], t=g[f'{tag}_t'][i]) for i in range(len(g[f'{tag}_R']))}


def _cam():
    return Camera(K=np.array([[30.0, 0, 16], [0, 30.0, 12], [0, 0, 1]]), dist=np.zeros(5))


@pytest.mark.parametrize('tag', ['plane7', 'tt12', 'tt40'])
def test_view_selection_matches_reference(tag):
    g = load('host_heuristics.npz')
    poses = _poses(g, tag)
    idxs = sorted(poses)
    pm = PatchMatchMVS(_cam(), scale=1.0)
    st = DenseStereoReconstructor(_cam(), scale=1.0)
    for k in (4, 5):
        for r in idxs:
            got = pm._select_source_views(r, idxs, poses, k=k)
            exp = [int(v) for v in g[f'{tag}_src_k{k}'][r] if v >= 0]
            assert got == exp, (tag, k, r)
    for r in idxs:
        exp = [int(v) for v in g[f'{tag}_nbr_k6'][r] if v >= 0]
        assert st._find_neighbors(r, idxs, poses, k=6) == exp


@pytest.mark.parametrize('tag', ['plane7', 'tt12', 'tt40'])
def test_depth_range_matches_reference(tag):
    g = load('host_heuristics.npz')
    poses = _poses(g, tag)
    pm = PatchMatchMVS(_cam(), scale=1.0)
    pm._estimate_depth_range(poses, None)
    assert [pm.depth_min, pm.depth_max] == list(g[f'{tag}_range_cams'])
    pm._estimate_depth_range(poses, g[f'{tag}_sparse'])
    assert [pm.depth_min, pm.depth_max] == list(g[f'{tag}_range_sparse'])


def test_intrinsics_scaling_matches_reference_conventions():
    K = np.array([[3144.93, 0, 1633.55], [0, 3148.59, 1990.96], [0, 0, 1]])
    cam = Camera(K=K, dist=np.zeros(5))
    pm = PatchMatchMVS(cam, scale=0.25)
    st = DenseStereoReconstructor(cam, scale=0.25)
    assert np.array_equal(pm.K_scaled[:2], K[:2] * 0.25) and pm.K_scaled[2, 2] == 1
    assert np.array_equal(st.K_scaled, pm.K_scaled)
    assert K[0, 0] == 3144.93                                 # caller's matrix untouched


def test_empty_results_for_fewer_than_three_cameras():
    poses = {0: CameraPose.identity(), 1: CameraPose.identity()}
    for obj in (PatchMatchMVS(_cam()), DenseStereoReconstructor(_cam())):
        p, c = obj.reconstruct([], poses)
        assert p.shape == (0,) and c.shape == (0,)


def test_filter_points_semantics():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(5000, 3))
    pts[:10] += 50.0                                          # far outliers
    col = rng.integers(0, 255, size=(5000, 3)).astype(np.uint8)
    p, c = PatchMatchMVS(_cam())._filter_points(pts, col)
    d = np.linalg.norm(pts - np.median(pts, axis=0), axis=1)
    keep = d < np.percentile(d, 95)
    v = np.floor(pts[keep] / 0.01).astype(np.int64)
    keys = v[:, 0] * 1000000000 + v[:, 1] * 1000000 + v[:, 2]
    _, first = np.unique(keys, return_index=True)
    assert np.array_equal(p, pts[keep][first]) and np.array_equal(c, col[keep][first])
    assert len(p) <= keep.sum() and np.abs(p).max() < 40


def test_voxel_down_sample_first_occurrence_sorted_by_key():
    pts = np.array([[0.011, 0.0, 0.0], [0.012, 0.001, 0.0], [-0.5, 0.2, 0.1], [0.011, 0.0, 0.0]])
    col = np.arange(12, dtype=np.uint8).reshape(4, 3)
    p, c = DenseStereoReconstructor(_cam())._voxel_down_sample(pts, col, 0.02)
    assert len(p) == 2 and np.array_equal(p[0], pts[2]) and np.array_equal(p[1], pts[0])
    e, _ = DenseStereoReconstructor(_cam())._voxel_down_sample(np.zeros((0, 3)), np.zeros((0, 3)), 0.02)
    assert len(e) == 0


def test_save_ply_format(tmp_path):
    pts = np.array([[1.0, -2.5, 3.1234567], [0.0, 1e-7, -1e3]])
    col = np.array([[255, 0, 7], [1, 2, 3]], dtype=np.uint8)
    out = tmp_path / 'sub' / 'x.ply'
    save_ply(pts, col, str(out))
    lines = out.read_text().splitlines()
    assert lines[:10] == ['ply', 'format ascii 1.0', 'element vertex 2', 'property float x', 'property float y',
                          'property float z', 'property uchar red', 'property uchar green', 'property uchar blue',
                          'end_header']
    exp = [f"{x:.6f} {y:.6f} {z:.6f} {int(r)} {int(g)} {int(b)}" for (x, y, z), (r, g, b) in zip(pts, col)]
    assert lines[10:] == exp
    save_ply(np.zeros((0, 3)), np.zeros((0, 3)), str(tmp_path / 'e.ply'))
    assert (tmp_path / 'e.ply').read_text().splitlines()[2] == 'element vertex 0'
    save_cameras_ply({0: CameraPose.identity()}, str(tmp_path / 'c.ply'), scale=0.5)
    cl = (tmp_path / 'c.ply').read_text().splitlines()
    assert cl[2] == 'element vertex 2' and cl[10].endswith('255 0 0') and cl[11].startswith('0.000000 0.000000 -0.500000')


def test_load_calibration_npz(tmp_path):
    f = tmp_path / 'calibration_data.npz'
    np.savez(f, mtx=np.eye(3, dtype=np.float32) * 2, dist=np.zeros((1, 3)))
    cam = load_calibration(str(f))
    assert cam.K.dtype == np.float64 and cam.dist.shape == (5,) and cam.fx == 2.0
    with pytest.raises(FileNotFoundError):
        load_calibration(str(tmp_path / 'nope.npz'))


def test_cli_flags_match_reference():
    import run_reconstruction as rr
    a = rr.build_parser().parse_args(['room', '--neural', '--mvs', '--max-images', '12', '--output', '/tmp/o'])
    assert (a.dataset, a.neural, a.mvs, a.stereo, a.dense, a.fast, a.max_images, a.output) == \
        ('room', True, True, False, False, False, 12, '/tmp/o')
    for flag in ('--dense', '--stereo', '--mvs', '--combined', '--fast', '--neural'):
        assert getattr(rr.build_parser().parse_args(['d', flag]), flag[2:]) is True


def test_depthnormalmap_fields():
    z = np.zeros((2, 2))
    dm = DepthNormalMap(depth=z, normal=np.zeros((2, 2, 3)), confidence=z)
    assert dm.depth is z and dm.normal.shape == (2, 2, 3)


def test_native_ply_writer_bytes_equal_reference_format(tmp_path):
    """b200d_write_ply_ascii writes byte-for-byte what the reference's per-point Python loop writes
    (utils.py:20-35), including negative zero, large and tiny magnitudes."""
    import time
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.normal(size=(20000, 3)) * 10, np.array([[-0.0, 1e-9, -1e-7], [123456.7891234, -5e5, 0.5]])])
    col = rng.integers(0, 256, size=(len(pts), 3)).astype(np.uint8)
    out = tmp_path / 'n.ply'
    t0 = time.perf_counter()
    save_ply(pts, col, str(out))
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    lines = ["ply", "format ascii 1.0", f"element vertex {len(pts)}", "property float x", "property float y",
             "property float z", "property uchar red", "property uchar green", "property uchar blue", "end_header"]
    for i in range(len(pts)):                                     # the reference's loop
        x, y, z = pts[i]
        r, g, b = col[i].astype(int)
        lines.append(f"{x:.6f} {y:.6f} {z:.6f} {r} {g} {b}")
    t_ref = time.perf_counter() - t0
    assert out.read_text() == "\n".join(lines) + "\n"
    assert t_native < t_ref


@pytest.mark.parametrize('n', [0, 1, 63, 64, 65, 100003])
def test_parallel_ply_writer_is_order_and_byte_exact(tmp_path, n):
    """The writer formats runs of vertices on all host threads and writes the runs in order: for any count (fewer
    vertices than threads, ragged last run, several 4M-vertex rounds are covered by the run arithmetic) the file equals
    the sequential formatting, including huge magnitudes that overflow the per-vertex reservation, inf and NaN."""
    rng = np.random.default_rng(n)
    pts = rng.normal(size=(n, 3)) * rng.choice([1e-3, 1.0, 1e4], size=(n, 1))
    if n > 10:
        pts[3] = [1e300, -1e300, 1e-300]
        pts[4] = [np.inf, -np.inf, np.nan]
        pts[5] = [-0.0, 0.4999995, -0.0000005]
        # the writer's own exact "%.6f" (128-bit integer rounding): binary ties round to even, subnormals print as zero,
        # values it declines (>= 2^52, beyond 64 bits after scaling) go through snprintf
        pts[6] = [0.0078125, -0.0078125, 0.0234375]
        pts[7] = [2.5e-6, 3.5e-6, 999999.9999995]
        pts[8] = [4503599627370496.0, 1e12, 1.9e13]
        pts[9] = [5e-324, -2.2250738585072014e-308, 0.9999995]
        pts[10] = [2.0 ** -21, -(2.0 ** -20), 2.0 ** -75]
    col = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
    out = tmp_path / 'p.ply'
    save_ply(pts, col, str(out))
    body = out.read_bytes().split(b'end_header\n', 1)[1]
    exp = ''.join('%.6f %.6f %.6f %d %d %d\n' % (*p, *c) for p, c in zip(pts.tolist(), col.tolist())).encode()
    assert body == exp


def test_oracle_extension_definitions_on_ground_truth():
    """CPU sanity of the two EXTENSION definitions in oracle/dense_oracle.c (no dackey-wav counterpart): exact depth maps
    are geometrically consistent in the sources that see the point and propose (almost) themselves; corrupted depths
    are rejected / corrected by the proposals."""
    from b200dense.engine import kinv32
    from b200dense.scenes import make_scene
    from oracle import oracle as orc
    sc = make_scene('wavy', 5, 120, 90, seed=12)
    ref, src = 2, [1, 3, 0, 4]
    R = [sc.poses[i][0] for i in range(5)]
    t = [sc.poses[i][1] for i in range(5)]
    cams = orc.Cams(sc.K, kinv32(sc.K), R[ref], t[ref], [R[i] for i in src], [t[i] for i in src])
    d_ref = sc.depth_gt[ref].copy()
    d_ref[30:50, 40:80] *= 1.2
    cnt = orc.geo_consistency(d_ref, [sc.depth_gt[i] for i in src], cams, 1.0, 0.01)
    assert cnt[30:50, 40:80].max() == 0 and (cnt[5:25, 40:80] >= 3).mean() > 0.9
    assert set(np.unique(cnt)) <= {0.0, 1.0, 2.0, 3.0, 4.0}
    H, W = d_ref.shape
    n_ref = np.zeros((H, W, 3), np.float32)
    n_ref[..., 2] = -1
    n_src = np.zeros((H, W, 3), np.float32)
    n_src[..., 0] = 1
    dmin, dmax = sc.depth_range
    cd, cn, allow = orc.viewprop_candidates(d_ref, n_ref, sc.depth_gt[src[0]], n_src, 0, cams, dmin, dmax)
    inner = allow[30:50, 40:80]
    assert inner.mean() > 0.9
    # the proposal pulls the corrupted block back towards the true surface (nearest-pixel lookup: not exactly)
    err_before = np.abs(d_ref - sc.depth_gt[ref])[30:50, 40:80][inner]
    err_after = np.abs(cd - sc.depth_gt[ref])[30:50, 40:80][inner]
    assert np.median(err_after) < 0.2 * np.median(err_before)
    # proposed normal = R_ref R_s^T n_s: unit length, and equal to the rotated source normal
    exp_n = (R[ref] @ R[src[0]].T @ np.array([1.0, 0, 0])).astype(np.float32)
    assert np.allclose(cn[allow], exp_n, atol=1e-6)
    # pixels without a proposal keep their hypothesis
    assert np.array_equal(cd[~allow], d_ref[~allow]) and np.array_equal(cn[~allow], n_ref[~allow])
