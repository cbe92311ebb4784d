"""Host->device upload of 4K photographs (24.9 MB uint8 each): driver's pageable path vs a page-locked staging ring
(memcpy on 1-8 threads into pooled pinned slots, then an asynchronous DMA), on the same box.  Prints host seconds (what
the enqueueing thread pays) and seconds until the data is on the device.

Measured on a B200 host (16 cores), idle GPU:   pageable .to(device) 10.9 GB/s | ring 1 thread 9.0 | 2 threads 16.3 |
4 threads 28.2 | page-locking a fresh copy 0.7 GB/s | plain memcpy 11.3 GB/s.
NOT ADOPTED: inside reconstruct() (200 x 4K photographs, 4 views in flight) the ring changed nothing measurable --
100-view jobs on one box: ring 37.7 / 28.0 views/s, direct 37.2 / 35.7 (profiles/microbench/r02s3_upload2_cmd.sh); the job
is bound by the kernels (39 views/s), the host thread only has to stay ahead of them, and the run-to-run spread of the
host side is larger than any difference between the two paths.  Routing the batch API's already page-locked inputs
through the ring cost 30 % (e2e 173 -> 121 views/s) before they were excluded."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-reconstruction-tool_b200')]
import numpy as np
import torch


class UploadRing:
    SLOTS = 4

    def __init__(self):
        from concurrent.futures import ThreadPoolExecutor
        self._slots, self._events, self._next = [None] * self.SLOTS, [None] * self.SLOTS, 0
        self._pool, self._threads = ThreadPoolExecutor(max_workers=8), 4

    def _executor(self):
        return self._pool

    def upload(self, a, tdtype, device):
        k = self._next % self.SLOTS
        self._next += 1
        if self._events[k] is not None:
            self._events[k].synchronize()
        n = a.nbytes
        buf = self._slots[k]
        if buf is None or buf.numel() < n:
            buf = self._slots[k] = torch.empty(max(n, 1 << 25), dtype=torch.uint8, pin_memory=True)
        dst, src = buf[:n].numpy(), a.reshape(-1).view(np.uint8)
        if self._threads == 1:
            np.copyto(dst, src)
        else:
            step = -(-n // self._threads)
            list(self._pool.map(lambda i: np.copyto(dst[i:i + step], src[i:i + step]), range(0, n, step)))
        out = torch.empty(a.shape, dtype=tdtype, device=device)
        out.copy_(buf[:n].view(tdtype).view(a.shape), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self._events[k] = ev
        return out


dev = torch.device('cuda:0')
N = 40
rng = np.random.default_rng(0)
imgs = [rng.integers(0, 255, size=(2160, 3840, 3), dtype=np.uint8) for _ in range(N)]
gb = sum(a.nbytes for a in imgs) / 1e9
torch.zeros(1, device=dev)
st = torch.cuda.Stream(device=dev)


def run(name, fn):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            keep = [fn(a) for a in imgs]
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        del keep
    print(f'{name:28s} host {t1 - t0:6.3f} s ({gb / (t1 - t0):5.1f} GB/s)   done {t2 - t0:6.3f} s ({gb / (t2 - t0):5.1f} GB/s)')


run('pageable .to(device)', lambda a: torch.from_numpy(a).to(dev, non_blocking=True))
for th in (1, 2, 4, 8):
    ring = UploadRing()
    ring._executor()
    ring._threads = th
    run(f'ring, {th} memcpy thread(s)', lambda a: ring.upload(a, torch.uint8, dev))
t0 = time.perf_counter()
pinned = [torch.from_numpy(a).pin_memory() for a in imgs[:8]]
print(f'page-locking copies: {sum(p.numel() for p in pinned) / 1e9 / (time.perf_counter() - t0):.1f} GB/s')
t0 = time.perf_counter(); dst = np.empty_like(imgs[0]); 
for a in imgs: np.copyto(dst, a)
print(f'plain memcpy pageable->pageable, 1 thread: {gb / (time.perf_counter() - t0):.1f} GB/s; cores {len(os.sched_getaffinity(0))}')
