"""GPU (-m gpu), needs >= 2 GPUs: a 2-rank NCCL `reconstruct()` returns on rank 0 exactly the cloud of a
1-rank run (same points, same order, same colours).  Skipped on a single-GPU box; the CPU/gloo test
(test_dist_gloo.py) covers the same plumbing without CUDA."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(rank, ws, port, q, mode='photometric'):
    for p in (ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200'), os.path.join(ROOT, '3d-reconstruction-tool_b200', 'src')):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from b200dense.scenes import make_scene
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    torch.cuda.set_device(rank)
    if ws > 1:
        dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=ws,
                                device_id=torch.device(f'cuda:{rank}'))
    sc = make_scene('wavy', 6, 128, 96, seed=21)
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in sc.poses}
    pm = PatchMatchMVS(cam, scale=1.0, patch_size=7, num_iterations=2, num_samples=2, min_views=2, seed=5,
                       device=f'cuda:{rank}', consistency='geometric' if mode == 'geometric' else 'photometric',
                       geo_min_views=1, geo_rel_thresh=0.02, geo_px_thresh=2.0,
                       schedule='viewprop' if mode == 'viewprop' else 'reference')
    sparse = np.stack([np.zeros(20), np.zeros(20), np.linspace(4.0, 6.0, 20)], axis=1)
    if mode == 'stereo':
        from core.dense_stereo import DenseStereoReconstructor
        ds = DenseStereoReconstructor(cam, scale=1.0, num_depths=24, patch_size=5, device=f'cuda:{rank}')
        p, c = ds.reconstruct(sc.images, poses)
        p, c = np.asarray(p).reshape(-1, 3), np.asarray(c).reshape(-1, 3)
    else:
        p, c = pm.reconstruct(sc.images, poses, sparse_points=sparse)
    q.put((rank, p, c))
    if ws > 1:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize('mode', ['photometric', 'geometric', 'viewprop', 'stereo'])
def test_two_rank_reconstruct_equals_single_rank(mode):
    """'geometric' additionally exercises the NCCL all-gather of the depth maps, 'viewprop' the per-iteration all-gather of
    depth AND normal maps (the proposals are evaluated from snapshots, so the result must not depend on the sharding),
    'stereo' the plane-sweep class.  With NVLink peer access the clouds travel as peer stores of the back-projection
    kernels (dist.peer_cloud), otherwise over NCCL point-to-point; the result must be the single-rank cloud either way."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    ctx = mp.get_context('spawn')
    out = {}
    for ws in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_run, args=(r, ws, port, q, mode)) for r in range(ws)]
        for p in procs:
            p.start()
        res = dict((r, (p, c)) for r, p, c in [q.get(timeout=240) for _ in range(ws)])
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        out[ws] = res
    p1, c1 = out[1][0]
    p2, c2 = out[2][0]
    assert len(p1) > 100
    assert np.array_equal(p1, p2) and np.array_equal(c1, c2)
    assert out[2][1][0].shape == (0, 3)
