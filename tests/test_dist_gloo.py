"""CPU, world_size 2, gloo: the multi-GPU plumbing (view sharding, ordered point gather, map all-gather).
The same code runs over NCCL on the GPU box; nothing here needs CUDA."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_cloud(view, n):
    g = np.random.default_rng(view)
    return (torch.from_numpy(g.normal(size=(n, 3))), torch.from_numpy(g.integers(0, 255, size=(n, 3)).astype(np.uint8)))


VIEWS = [0, 1, 2, 3, 5, 8, 13]                       # 7 views: uneven shards
COUNTS = {0: 5, 1: 0, 2: 17, 3: 1, 5: 9, 8: 0, 13: 4}  # some views fuse to nothing


def _worker(rank, ws, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, '3d-reconstruction-tool_b200'))
    from b200dense import dist as bdist
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=ws)
    try:
        mine = bdist.shard_views(VIEWS)
        my_pos = list(bdist.shard_positions(len(VIEWS)))
        blocks = {2: [[0, 1, 2, 3], [4, 5, 6]], 3: [[0, 1, 2], [3, 4], [5, 6]]}[ws]                          # contiguous
        assert mine == [VIEWS[p] for p in my_pos] and my_pos == blocks[rank]
        local = {v: _fake_cloud(v, COUNTS[v]) for v in mine if v != 8}     # view 8 missing entirely
        pts, col = bdist.gather_point_clouds(VIEWS, local, torch.device('cpu'))
        maps = bdist.all_gather_maps(VIEWS, {v: torch.full((3, 4), float(v)) for v in mine}, (3, 4), torch.device('cpu'))
        ok_maps = all(torch.equal(maps[v], torch.full((3, 4), float(v))) for v in VIEWS)
        # the copy-free gather reconstruct() uses: counts exchanged once, rank 0 receives straight into final slices
        counts = bdist.exchange_counts(len(VIEWS), {p: COUNTS[VIEWS[p]] for p in my_pos}, torch.device('cpu'))
        assert counts == [COUNTS[v] for v in VIEWS]
        offs = np.concatenate([[0], np.cumsum(counts)])
        total = int(offs[-1])
        if rank == 0:
            out_p = torch.full((total, 3), -7.0, dtype=torch.float64)
            out_c = torch.zeros((total, 3), dtype=torch.uint8)
            loc = {}
            for p in my_pos:                                     # own views are written in place
                cp, cc = _fake_cloud(VIEWS[p], COUNTS[VIEWS[p]])
                out_p[offs[p]:offs[p + 1]] = cp
                out_c[offs[p]:offs[p + 1]] = cc
                loc[p] = (out_p[offs[p]:offs[p + 1]], out_c[offs[p]:offs[p + 1]])
            bdist.gather_into(counts, loc, out_p, out_c)
            ok_maps = ok_maps and torch.equal(out_p, pts) and torch.equal(out_c, col)
        else:
            loc = {p: _fake_cloud(VIEWS[p], COUNTS[VIEWS[p]]) for p in my_pos}
            bdist.gather_into(counts, loc, None, None)
        q.put((rank, pts.numpy(), col.numpy(), ok_maps))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize('ws', [2, 3])
def test_gather_point_clouds_and_maps_multi_rank(ws):
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(ws):
        r, pts, col, ok = q.get(timeout=90)
        res[r] = (pts, col, ok)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    exp_p = np.concatenate([_fake_cloud(v, COUNTS[v])[0].numpy() for v in VIEWS])
    exp_c = np.concatenate([_fake_cloud(v, COUNTS[v])[1].numpy() for v in VIEWS])
    assert np.array_equal(res[0][0], exp_p) and np.array_equal(res[0][1], exp_c)     # sorted-view order on rank 0
    for r in range(1, ws):
        assert res[r][0].shape == (0, 3) and res[r][1].shape == (0, 3)
    assert all(res[r][2] for r in range(ws))


def test_single_process_paths():
    import sys
    from b200dense import dist as bdist
    assert bdist.world() == (0, 1)
    assert bdist.shard_views([4, 5, 6], 1, 2) == [6] and bdist.shard_views([4, 5, 6], 0, 2) == [4, 5]
    local = {v: _fake_cloud(v, COUNTS[v]) for v in VIEWS}
    pts, col = bdist.gather_point_clouds(VIEWS, local, torch.device('cpu'))
    assert pts.shape[0] == sum(COUNTS.values()) and col.dtype == torch.uint8
    e, _ = bdist.gather_point_clouds(VIEWS, {}, torch.device('cpu'))
    assert e.shape == (0, 3)
    maps = bdist.all_gather_maps([7, 9], {7: torch.ones(2, 2), 9: torch.zeros(2, 2)}, (2, 2), torch.device('cpu'))
    assert torch.equal(maps[7], torch.ones(2, 2)) and torch.equal(maps[9], torch.zeros(2, 2))
    # single rank: counts are the local ones, the in-place gather is a no-op, and there is no peer buffer to map
    assert bdist.exchange_counts(4, {0: 3, 2: 5}, torch.device('cpu')) == [3, 0, 5, 0]
    bdist.gather_into([3, 0, 5, 0], {}, None, None)
    assert bdist.peer_cloud(torch.device('cpu'), 100) is None
    # contiguous blocks whose sizes differ by at most one; owner_of is the inverse of shard_positions for any (n, world)
    for n in (0, 1, 3, 7, 8, 9, 64, 200):
        for ws in (1, 2, 3, 4, 8):
            blocks = [list(bdist.shard_positions(n, r, ws)) for r in range(ws)]
            assert sum(blocks, []) == list(range(n))
            assert max(map(len, blocks)) - min(map(len, blocks)) <= 1
            assert all(bdist.owner_of(p, ws, n) == r for r in range(ws) for p in blocks[r])
