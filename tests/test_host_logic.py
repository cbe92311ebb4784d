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
