"""GPU (-m gpu): the entry point with the reference's flags (src/run_reconstruction.py:24-186) end to end on a synthetic
dataset with known poses: images on disk + calibration npz + poses npz -> dense_mvs.ply / dense_stereo.ply, the stdout
keywords the reference's GUI greps (gui/worker.py:112-122), and the same cloud as calling the class directly."""
import contextlib
import io
import os

import cv2 as cv
import numpy as np
import pytest

from b200dense.scenes import make_scene
from core.camera import Camera, CameraPose

pytestmark = pytest.mark.gpu


def _dataset(tmp_path):
    sc = make_scene('wavy', 6, 256, 192, seed=13, arc_deg=5.0)
    img_dir = tmp_path / 'scene'
    img_dir.mkdir()
    files = []
    for i, im in enumerate(sc.images):
        name = f'img_{i:02d}.png'
        cv.imwrite(str(img_dir / name), im['image'])
        files.append(name)
    np.savez(tmp_path / 'calibration_data.npz', mtx=sc.K, dist=np.zeros((1, 5)))
    sparse = np.stack([np.zeros(30), np.zeros(30), np.linspace(4.0, 6.0, 30)], axis=1)
    np.savez(tmp_path / 'poses.npz', R=np.stack([sc.poses[i][0] for i in range(6)]),
             t=np.stack([sc.poses[i][1] for i in range(6)]), files=np.array(files), points=sparse)
    return sc, img_dir, sparse


def _read_ply(path):
    lines = open(path).read().splitlines()
    n = int(lines[2].split()[-1])
    body = np.array([ln.split() for ln in lines[10:10 + n]], dtype=np.float64).reshape(n, 6)
    return body[:, :3], body[:, 3:].astype(np.uint8)


@pytest.mark.parametrize('flag,ply', [('--mvs', 'dense_mvs.ply'), ('--stereo', 'dense_stereo.ply')])
def test_cli_known_poses_writes_the_dense_cloud(tmp_path, flag, ply):
    import run_reconstruction as rr
    sc, img_dir, sparse = _dataset(tmp_path)
    out_dir = tmp_path / 'out'
    buf = io.StringIO()
    np.random.seed(0)
    with contextlib.redirect_stdout(buf):
        rr.main([str(img_dir), flag, '--poses', str(tmp_path / 'poses.npz'), '--calibration',
                 str(tmp_path / 'calibration_data.npz'), '--output', str(out_dir), '--scale', '0.5'])
    text = buf.getvalue()
    assert os.path.exists(out_dir / ply) and os.path.exists(out_dir / 'sparse.ply') and os.path.exists(out_dir / 'cameras.ply')
    low = text.lower()
    assert ('dense' in low or 'stereo' in low) and 'saved' in low and '.ply' in low and 'DONE!' in text
    pts, col = _read_ply(out_dir / ply)
    assert len(pts) > 100
    # the same cloud as the class called directly with what the CLI passes (PNG round trip is lossless)
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    images = [{'image': cv.imread(str(img_dir / f'img_{i:02d}.png'))} for i in range(6)]
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in range(6)}
    with contextlib.redirect_stdout(io.StringIO()):
        if flag == '--mvs':
            from core.mvs_patchmatch import PatchMatchMVS
            p, c = PatchMatchMVS(cam, scale=0.5, num_iterations=3, min_views=3).reconstruct(images, poses, sparse_points=sparse)
        else:
            from core.dense_stereo import DenseStereoReconstructor
            np.random.seed(0)
            p, c = DenseStereoReconstructor(cam, scale=0.5).reconstruct(images, poses, max_pairs=30)
    assert len(p) == len(pts) and np.array_equal(c, col)
    assert np.abs(p - pts).max() <= 5.1e-7                            # "%.6f"
