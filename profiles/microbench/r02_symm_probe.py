"""Probe: is torch symmetric memory (peer buffers over NVLink) usable on this box?  torchrun --nproc-per-node 2 this file.
Rank 1 writes a pattern straight into rank 0's buffer with a copy kernel; rank 0 checks it after a barrier."""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
try:
    n = 64 << 20
    t0 = time.perf_counter()
    buf = symm_mem.empty(n, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    buf.zero_()
    hdl.barrier()
    if rank != 0:
        peer0 = hdl.get_buffer(0, (n,), torch.uint8)
        src = torch.full((1 << 20,), rank, dtype=torch.uint8, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        big = torch.full((n // 2,), 7, dtype=torch.uint8, device=dev)
        e0.record()
        peer0[rank << 20:(rank + 1) << 20].copy_(src)
        peer0[n // 2:].copy_(big)
        e1.record()
        torch.cuda.synchronize()
        print(f'rank {rank}: wrote into rank 0, {n // 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s', flush=True)
    hdl.barrier()
    if rank == 0:
        torch.cuda.synchronize()
        ok = all(int(buf[r << 20].item()) == r and int(buf[((r + 1) << 20) - 1].item()) == r for r in range(1, world))
        ok = ok and int(buf[n - 1].item()) == 7
        print(f'symmetric memory ok={ok}; alloc+rendezvous {1e3 * (t1 - t0):.1f} ms', flush=True)
except Exception as e:  # noqa: BLE001
    print(f'rank {rank}: symmetric memory unavailable: {type(e).__name__}: {e}', flush=True)
dist.destroy_process_group()
