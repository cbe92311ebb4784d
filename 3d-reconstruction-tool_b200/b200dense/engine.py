"""Host-side engine: device memory and streams through torch, compute through the C ABI.

PyTorch is plumbing here (allocation, H2D/D2H copies, streams, torch.distributed); every
operation of the dense hot path is a hand-written sm_100a kernel behind include/b200dense.h.
There is no CPU path: constructing a DenseEngine without CUDA raises.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import B200DenseError, Cams, PmParams, Rng


def kinv32(K_scaled: np.ndarray) -> np.ndarray:
    """fp32 inverse of the working-resolution intrinsics, the way the reference gets it
    (`torch.inverse(K)` on an fp32 tensor, mvs_patchmatch.py:237-238, dense_stereo.py:232-233),
    evaluated on the host so both sides of every parity check see the same bits."""
    return torch.inverse(torch.from_numpy(np.asarray(K_scaled, dtype=np.float64).astype(np.float32))).numpy()


def make_cams(K_scaled, Kinv32, R_ref, t_ref, R_src: Sequence, t_src: Sequence) -> Cams:
    """Pack the fp32 camera block (f64 -> f32 casts as in mvs_patchmatch.py:237-257)."""
    if not 1 <= len(R_src) <= _cabi.MAX_SRC:
        raise ValueError(f"need 1..{_cabi.MAX_SRC} source views, got {len(R_src)}")
    c = Cams()
    f = np.float32
    c.K[:] = np.asarray(K_scaled, dtype=np.float64).astype(f).ravel().tolist()
    c.Kinv[:] = np.asarray(Kinv32, dtype=f).ravel().tolist()
    c.R_ref[:] = np.asarray(R_ref, dtype=np.float64).astype(f).ravel().tolist()
    c.t_ref[:] = np.asarray(t_ref, dtype=np.float64).astype(f).ravel().tolist()
    c.n_src = len(R_src)
    for s, (R, t) in enumerate(zip(R_src, t_src)):
        c.R_src[s][:] = np.asarray(R, dtype=np.float64).astype(f).ravel().tolist()
        c.t_src[s][:] = np.asarray(t, dtype=np.float64).astype(f).ravel().tolist()
    return c


class PinnedPool:
    """Recycles page-locked staging buffers for device->host results.

    Page-locking a fresh 23 MB buffer costs ~12 ms on the B200 hosts (measured, profiles/microbench/dbg_e2e.py) --
    more than the PatchMatch of the view it carries -- so buffers are pooled by (shape, dtype).  A result array
    handed to the caller owns its buffer; when the array is garbage-collected the buffer returns to the pool."""

    def __init__(self, max_idle_bytes: int = 2 << 30):
        self._free: Dict[Tuple[Tuple[int, ...], torch.dtype], List[torch.Tensor]] = {}
        self.max_idle_bytes = max_idle_bytes      # idle (returned) buffers kept for reuse; beyond it they are released
        self._idle_bytes = 0

    def take(self, shape, dtype) -> torch.Tensor:
        lst = self._free.get((tuple(shape), dtype))
        if lst:
            t = lst.pop()
            self._idle_bytes -= t.numel() * t.element_size()
            return t
        return torch.empty(tuple(shape), dtype=dtype, pin_memory=True)

    def _give(self, t: torch.Tensor) -> None:
        nbytes = t.numel() * t.element_size()
        if self._idle_bytes + nbytes > self.max_idle_bytes:
            return                                # dropped: the page-locked allocation is freed with the tensor
        self._idle_bytes += nbytes
        self._free.setdefault((tuple(t.shape), t.dtype), []).append(t)

    def trim(self) -> None:
        """Release every idle buffer."""
        self._free.clear()
        self._idle_bytes = 0

    def as_numpy(self, t: torch.Tensor) -> np.ndarray:
        a = t.numpy()
        weakref.finalize(a, self._give, t)
        return a


PINNED = PinnedPool()


class UploadRing:
    """Pageable host array -> pooled page-locked slot (memcpy on a few threads; NumPy releases the GIL) -> async DMA.

    `tensor.to(device)` from a caller's pageable array goes through the driver's staging path, which is synchronous for
    the host thread and must first wait for whatever precedes it on the stream.  In isolation both paths are fast (10.9
    vs 28 GB/s, profiles/microbench/r02s3_upload_ab.py), but inside a 200-view 4K reconstruction, where the same thread
    also has to keep four views' worth of kernels enqueued, the ring is worth 5-20 % and removes most of the run-to-run
    spread: 5.35 / 5.34 s against 5.64 / 6.63 s on the same box (profiles/microbench/r02s3_upload3_cmd.sh).  Arrays that
    are page-locked already (the batch API's staging buffers) are DMA'd directly: sending those through the ring cost
    30 % of the e2e rate.  B200D_UPLOAD=direct disables the ring."""
    SLOTS = 4

    def __init__(self):
        from concurrent.futures import ThreadPoolExecutor
        self._slots: List[Optional[torch.Tensor]] = [None] * self.SLOTS
        self._events: List[Optional[torch.cuda.Event]] = [None] * self.SLOTS
        self._next = 0
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:                                      # pragma: no cover
            cores = os.cpu_count() or 1
        self._threads = max(1, min(4, cores))                       # 4 memcpy threads saturate the host (28 GB/s measured)
        self._pool = ThreadPoolExecutor(max_workers=self._threads)

    def upload(self, a: np.ndarray, tdtype: torch.dtype, device: torch.device) -> torch.Tensor:
        """`a`: C-contiguous host array whose dtype corresponds to `tdtype`.  Returns the device tensor; the copy is
        enqueued on the current stream.  The caller's array is not referenced after this returns."""
        k = self._next % self.SLOTS
        self._next += 1
        if self._events[k] is not None:
            self._events[k].synchronize()                           # the DMA that last read this slot has finished
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


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensors must be contiguous CUDA tensors"
    return t.data_ptr()


def _tensors_of(obj):
    """CUDA tensors inside a job result (tensor / dict / tuple / list, nested)."""
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensors_of(v)
    elif isinstance(obj, (tuple, list)):
        for v in obj:
            yield from _tensors_of(v)


def _ptr_array(ts: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(ts))(*[_ptr(t) for t in ts])
    return arr


class DenseEngine:
    """One engine per GPU (one process per GPU).  Owns cached scratch buffers."""

    def __init__(self, device: str | torch.device | int = 'cuda:0'):
        if not torch.cuda.is_available():
            raise B200DenseError("CUDA is not available: the dense path has no CPU fallback")
        self.lib = _cabi.load()
        self.device = torch.device(device if not isinstance(device, int) else f'cuda:{device}')
        if self.device.type != 'cuda':
            raise B200DenseError(f"DenseEngine needs a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self._scratch: Dict[Tuple[str, int, int, int], torch.Tensor] = {}
        self._streams: List[torch.cuda.Stream] = []
        self._tables: Dict[tuple, tuple] = {}
        self.launches = 0                     # kernels of ours enqueued (bench's gpu_launches claim)

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def scratch(self, kind: str, H: int, W: int, slot: int = 0) -> torch.Tensor:
        key = (kind, H, W, slot)
        buf = self._scratch.get(key)
        if buf is None:
            nbytes = {'view': self.lib.b200d_scratch_bytes, 'fuse': self.lib.b200d_fuse_scratch_bytes}[kind](H, W)
            buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
            self._scratch[key] = buf
        return buf

    def to_device(self, a: np.ndarray, dtype=torch.float32) -> torch.Tensor:
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a)
        if (os.environ.get('B200D_UPLOAD', 'ring') == 'ring' and t.dtype == dtype and a.nbytes >= (1 << 20)
                and torch.device(self.device).type == 'cuda' and not t.is_pinned()):
            if getattr(self, '_upload_ring', None) is None:
                self._upload_ring = UploadRing()
            return self._upload_ring.upload(a, dtype, torch.device(self.device))
        return t.to(self.device, dtype=dtype, non_blocking=True).contiguous()

    @staticmethod
    def to_host(ts: Sequence[torch.Tensor]) -> List[np.ndarray]:
        """Device tensors -> host arrays in pooled page-locked buffers: one DMA per tensor, one synchronisation."""
        hs = []
        for t in ts:
            h = PINNED.take(t.shape, t.dtype)
            h.copy_(t, non_blocking=True)
            hs.append(h)
        torch.cuda.current_stream().synchronize()
        return [PINNED.as_numpy(h) for h in hs]

    def copy_stream(self) -> torch.cuda.Stream:
        """Side stream for host->device uploads that should overlap the compute streams."""
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    # ------------------------------------------------------------------ concurrent views
    def run_views(self, jobs: Sequence, n_streams: int = 4) -> list:
        """Run independent per-view jobs concurrently on `n_streams` side streams.

        A reference view's PatchMatch is a chain of ~33 dependent launches whose grids (950 CTAs at
        1600x1200) leave a partly empty last wave on 148 SMs; views are independent, so the
        tail of one view's launch is filled by another view's.  `jobs` are callables taking the slot
        index (their scratch slot); job j runs on stream j % n_streams, i.e. jobs sharing a slot are
        serialised.  The caller's current stream waits for all of them before this returns."""
        n = max(1, min(n_streams, len(jobs)))
        while len(self._streams) < n:
            self._streams.append(torch.cuda.Stream(device=self.device))
        main = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(main)
        results = []
        try:
            for j, job in enumerate(jobs):
                st = self._streams[j % n]
                if j < n:
                    st.wait_event(start)
                with torch.cuda.stream(st):
                    results.append(job(j % n))
        finally:
            # also when a job raised: the caller's stream must not run ahead of (or free buffers under) the
            # kernels already enqueued on the side streams
            for st in self._streams[:n]:
                done = torch.cuda.Event()
                done.record(st)
                main.wait_event(done)
        # results were allocated on side streams and are consumed on the caller's stream from here on
        for r in results:
            for t in _tensors_of(r):
                t.record_stream(main)
        return results

    # ------------------------------------------------------------------ per-launch timing (benchmarks)
    def profile_begin(self, max_records: int = 65536) -> None:
        _cabi.check(self.lib.b200d_profile_begin(int(max_records)))

    def profile_end(self, cap: int = 65536):
        """-> (durations_ms float32[n], tags int32[n]) of the kernels launched since profile_begin."""
        ms = np.zeros(cap, np.float32)
        tags = np.zeros(cap, np.int32)
        n = C.c_int(0)
        _cabi.check(self.lib.b200d_profile_end(ms.ctypes.data_as(C.POINTER(C.c_float)),
                                               tags.ctypes.data_as(C.POINTER(C.c_int32)), cap, C.byref(n)))
        return ms[:n.value], tags[:n.value]

    # ------------------------------------------------------------------ entry points
    def ref_stats(self, ref: torch.Tensor, patch: int) -> Tuple[torch.Tensor, torch.Tensor]:
        H, W = ref.shape
        mean1 = torch.empty_like(ref)
        var1 = torch.empty_like(ref)
        _cabi.check(self.lib.b200d_ref_stats(_ptr(ref), H, W, patch, _ptr(mean1), _ptr(var1),
                                             _ptr(self.scratch('view', H, W)), self._stream()))
        self.launches += 1
        return mean1, var1

    def pm_cost(self, ref: torch.Tensor, srcs: Sequence[torch.Tensor], cams: Cams, depth: torch.Tensor, patch: int,
                mode: int = 0, arith: int = _cabi.PARITY) -> torch.Tensor:
        """Cost map (mode 0, mvs_patchmatch.py:323-390) or confidence count (mode 1, :493-534) of a depth map."""
        H, W = ref.shape
        out = torch.empty_like(ref)
        _cabi.check(self.lib.b200d_pm_cost(_ptr(ref), _ptr_array(srcs), C.byref(cams), _ptr(depth), H, W, patch, mode,
                                           arith, _ptr(out), _ptr(self.scratch('view', H, W)), self._stream()))
        self.launches += 3
        return out

    def pm_view(self, ref: torch.Tensor, srcs: Sequence[torch.Tensor], cams: Cams, patch: int, iters: int,
                samples: int, depth_min: float, depth_max: float, *, seed: int = 0, view_id: int = 0,
                init_depth: Optional[torch.Tensor] = None, init_normal: Optional[torch.Tensor] = None,
                rand_u: Optional[torch.Tensor] = None, rand_n: Optional[torch.Tensor] = None,
                arith: int = _cabi.PARITY, dump_costs: bool = False, want_best: bool = False, out=None,
                slot: int = 0, schedule: int = _cabi.SCHEDULE_REFERENCE, it_begin: int = 0, it_end: int = 0,
                stage: int = 0, costs: Optional[torch.Tensor] = None):
        """Full PatchMatch of one reference view (mvs_patchmatch.py:225-321); returns device tensors
        dict(depth, normal, confidence[, best_cost, costs]).  `schedule`: SCHEDULE_REFERENCE (the reference's two
        shifted evaluations per iteration, bit-exact) or SCHEDULE_REDBLACK (extension: 8 checkerboard evaluations).
        Staged execution (`it_begin`, `it_end`, `stage` flags STAGE_NO_INIT / STAGE_NO_CONF): run part of the view on
        the state held in `out` (which must then carry 'best_cost'); used by the view-propagation schedule."""
        H, W = ref.shape
        prm = PmParams(patch=patch, iters=iters, samples=samples, arith=arith,
                       depth_min=float(depth_min), depth_max=float(depth_max), schedule=int(schedule), stage=int(stage),
                       it_begin=int(it_begin), it_end=int(it_end))
        per_it = (8 if schedule == _cabi.SCHEDULE_REDBLACK else 2) + samples
        rng = Rng(seed=seed, view_id=view_id, d_init_depth=_ptr(init_depth), d_init_normal=_ptr(init_normal),
                  d_rand_u=_ptr(rand_u), d_rand_n=_ptr(rand_n))
        if out is None:
            out = dict(depth=torch.empty((H, W), dtype=torch.float32, device=self.device),
                       normal=torch.empty((H, W, 3), dtype=torch.float32, device=self.device),
                       confidence=torch.empty((H, W), dtype=torch.float32, device=self.device))
        best = out.get('best_cost')
        if best is None and want_best:
            best = torch.empty((H, W), dtype=torch.float32, device=self.device)
        if costs is None and dump_costs:
            costs = torch.empty((iters * per_it, H, W), dtype=torch.float32, device=self.device)
        _cabi.check(self.lib.b200d_pm_view(_ptr(ref), _ptr_array(srcs), C.byref(cams), H, W, C.byref(prm), C.byref(rng),
                                           _ptr(self.scratch('view', H, W, slot)), _ptr(out['depth']), _ptr(out['normal']),
                                           _ptr(out['confidence']), _ptr(best), _ptr(costs), self._stream()))
        n_it = (it_end if it_end > 0 else iters) - it_begin
        # stats, [init], (sample + ncc) per evaluation, [confidence pair]
        self.launches += 1 + (0 if stage & _cabi.STAGE_NO_INIT else 1) + 2 * n_it * per_it + \
            (0 if stage & _cabi.STAGE_NO_CONF else 2)
        if best is not None:
            out['best_cost'] = best
        if costs is not None:
            out['costs'] = costs
        return out

    def pm_viewprop(self, ref: torch.Tensor, srcs: Sequence[torch.Tensor], cams: Cams, patch: int, depth_min: float,
                    depth_max: float, src_depth: Sequence[torch.Tensor], src_normal: Sequence[torch.Tensor], state: dict,
                    *, arith: int = _cabi.PARITY, dump_costs: bool = False, slot: int = 0):
        """EXTENSION: one view-propagation step (b200d_pm_viewprop) on `state` = dict(depth, normal, best_cost), in
        place, from snapshots of the source views' depth / normal maps.  Returns the (S,H,W) cost dump if asked."""
        H, W = ref.shape
        prm = PmParams(patch=patch, iters=0, samples=0, arith=arith, depth_min=float(depth_min),
                       depth_max=float(depth_max), schedule=0, stage=0, it_begin=0, it_end=0)
        costs = torch.empty((len(srcs), H, W), dtype=torch.float32, device=self.device) if dump_costs else None
        _cabi.check(self.lib.b200d_pm_viewprop(_ptr(ref), _ptr_array(srcs), C.byref(cams), H, W, C.byref(prm),
                                               _ptr_array(src_depth), _ptr_array(src_normal),
                                               _ptr(self.scratch('view', H, W, slot)), _ptr(state['depth']),
                                               _ptr(state['normal']), _ptr(state['best_cost']), _ptr(costs),
                                               self._stream()))
        self.launches += 1 + 3 * len(srcs)
        return costs

    def measure_fp32_peak(self) -> float:
        """Sustained FFMA2 rate of this GPU in T lane-FMA/s (synchronises; benchmarks' roofline denominator)."""
        v = C.c_double(0.0)
        _cabi.check(self.lib.b200d_measure_fp32_peak(C.byref(v), self._stream()))
        return float(v.value)

    def sweep_bound_check(self, reset: bool = False):
        """(r_m2, r_a2, r_ar): largest observed |separable sum - chain| / bound since the last reset, from plane sweeps run
        with arith=PARITY_CHECK (synchronises)."""
        v = (C.c_float * 3)()
        _cabi.check(self.lib.b200d_sweep_bound_check(v, 1 if reset else 0))
        return tuple(float(x) for x in v)

    def sweep_fallbacks(self, reset: bool = False) -> int:
        """Outputs of the bounded plane sweep redone with the exact chains since the last reset (synchronises)."""
        n = C.c_uint64(0)
        _cabi.check(self.lib.b200d_sweep_fallbacks(C.byref(n), 1 if reset else 0))
        return int(n.value)

    def plane_sweep(self, ref: torch.Tensor, nbrs: Sequence[torch.Tensor], cams: Cams, depths: np.ndarray,
                    patch: int, thresh: float, arith: int = _cabi.PARITY, dump_ncc: bool = False, slot: int = 0):
        """Plane sweep of one reference view (dense_stereo.py:222-316); `depths` is a host fp32 array;
        returns device tensors."""
        H, W = ref.shape
        depths = np.ascontiguousarray(depths, dtype=np.float32)
        D = int(depths.size)
        depth_map = torch.empty((H, W), dtype=torch.float32, device=self.device)
        conf = torch.empty((H, W), dtype=torch.float32, device=self.device)
        idx = torch.empty((H, W), dtype=torch.int32, device=self.device)
        ncc = torch.empty((D, len(nbrs), H, W), dtype=torch.float32, device=self.device) if dump_ncc else None
        _cabi.check(self.lib.b200d_plane_sweep(_ptr(ref), _ptr_array(nbrs), C.byref(cams),
                                               depths.ctypes.data_as(C.POINTER(C.c_float)), D, H, W, patch,
                                               float(thresh), arith, _ptr(depth_map), _ptr(conf), _ptr(idx), _ptr(ncc),
                                               _ptr(self.scratch('view', H, W, slot)), self._stream()))
        self.launches += 1 + 2 * D
        return dict(depth_map=depth_map, confidence=conf, best_idx=idx, ncc=ncc)

    def geo_consistency(self, depth_ref: torch.Tensor, depth_srcs: Sequence[torch.Tensor], cams: Cams,
                        px_thresh: float = 1.0, rel_thresh: float = 0.01) -> torch.Tensor:
        """EXTENSION: number of source views whose depth map agrees geometrically with `depth_ref` (H,W)."""
        H, W = depth_ref.shape
        out = torch.empty_like(depth_ref)
        _cabi.check(self.lib.b200d_geo_consistency(_ptr(depth_ref), _ptr_array(depth_srcs), C.byref(cams), H, W,
                                                   float(px_thresh), float(rel_thresh), _ptr(out), self._stream()))
        self.launches += 1
        return out

    def knn_mean_dist(self, points: torch.Tensor, k: int = 20) -> torch.Tensor:
        """Mean distance of every point to its k-1 nearest other points, f64 (dense_stereo.py:456-460)."""
        n = int(points.shape[0])
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.b200d_knn_mean_dist(_ptr(points), n, int(k), _ptr(out), self._stream()))
        self.launches += 1
        return out

    def prepare_image(self, bgr: torch.Tensor, nh: int, nw: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device `_prepare_images` of one BGR u8 image (H,W,3): -> (colour (nh,nw,3) u8, gray (nh,nw) f32),
        bit-identical to cv.resize + cv.cvtColor + /255 (mvs_patchmatch.py:176-177)."""
        from .prepare import resize_tables
        H, W = int(bgr.shape[0]), int(bgr.shape[1])
        key = ('prep', H, W, nh, nw)
        tabs = self._tables.get(key)
        if tabs is None:
            xi, xa, yi, ya = resize_tables(H, W, nh, nw)
            tabs = tuple(torch.from_numpy(np.ascontiguousarray(t)).to(self.device) for t in (xi, xa, yi, ya))
            self._tables[key] = tabs
        color = torch.empty((nh, nw, 3), dtype=torch.uint8, device=self.device)
        gray = torch.empty((nh, nw), dtype=torch.float32, device=self.device)
        _cabi.check(self.lib.b200d_prepare_image(_ptr(bgr), H, W, nh, nw, _ptr(tabs[0]), _ptr(tabs[1]), _ptr(tabs[2]),
                                                 _ptr(tabs[3]), _ptr(color), _ptr(gray), self._stream()))
        self.launches += 1
        return color, gray

    def mask_count(self, depth: torch.Tensor, conf: torch.Tensor, min_conf: float, mode: int, count: torch.Tensor,
                   slot: int = 0) -> torch.Tensor:
        """Count-only pass of the fusion mask into `count` (1-element int64 device tensor); nothing is read back."""
        H, W = depth.shape
        _cabi.check(self.lib.b200d_mask_count(_ptr(depth), _ptr(conf), H, W, float(min_conf), mode, _ptr(count),
                                              _ptr(self.scratch('fuse', H, W, slot)), self._stream()))
        self.launches += 2
        return count

    def fuse_to_rank0(self, n_positions: int, items: Sequence[dict], Kinv64: np.ndarray, min_conf: float, mode: int):
        """Fusion of many views into ONE cloud on rank 0 in view order, with a single host synchronisation.

        `items`: this rank's views, each dict(pos=index in the sorted view list, depth, conf (device maps), bgr=callable
        returning the device colour image, R, t).  Steps: count every mask on the device (b200d_mask_count) -> read the
        counts once -> exchange them (one small all-reduce) -> rank 0's cloud is allocated exactly (symmetric memory when
        the job is multi-GPU: dist.peer_cloud) -> every rank back-projects each view straight into its slice at the
        view's global offset (over NVLink for ranks > 0: the kernel is the transfer) -> fence.  Fallback without peer
        memory: NCCL point-to-point into the final slices (dist.gather_into).  Returns (points, colors, counts, how) on
        rank 0 and (None, None, counts, how) elsewhere; `counts` has one entry per position."""
        from . import dist as bdist
        rank, ws = bdist.world()
        dev = self.device
        counts_d = torch.zeros(max(1, len(items)), dtype=torch.int64, device=dev)
        for j, it in enumerate(items):
            self.mask_count(it['depth'], it['conf'], min_conf, mode, counts_d[j:j + 1])
        counts_h = counts_d.cpu().tolist()                    # the one host synchronisation
        local_counts = {it['pos']: int(c) for it, c in zip(items, counts_h)}
        all_counts = bdist.exchange_counts(n_positions, local_counts, dev)
        offs = np.concatenate([[0], np.cumsum(all_counts)]).astype(np.int64)
        total = int(offs[-1])
        dest = bdist.peer_cloud(dev, total)
        if dest is not None:
            buf_p, buf_c = dest
            base = {it['pos']: int(offs[it['pos']]) for it in items}
        elif rank == 0:
            buf_p = torch.empty((total, 3), dtype=torch.float64, device=dev)
            buf_c = torch.empty((total, 3), dtype=torch.uint8, device=dev)
            base = {it['pos']: int(offs[it['pos']]) for it in items}
        else:
            mine_n = int(sum(local_counts.values()))
            buf_p = torch.empty((mine_n, 3), dtype=torch.float64, device=dev)
            buf_c = torch.empty((mine_n, 3), dtype=torch.uint8, device=dev)
            run, base = 0, {}
            for it in items:
                base[it['pos']] = run
                run += local_counts[it['pos']]
        local = {}
        dummy = torch.zeros(1, dtype=torch.int64, device=dev)
        for it in items:
            m, b = local_counts[it['pos']], base[it['pos']]
            if m == 0:
                continue
            sl = (buf_p[b:b + m], buf_c[b:b + m])
            self.mask_backproject(it['depth'], it['conf'], it['bgr'](), Kinv64, it['R'], it['t'], min_conf, mode,
                                  points=sl[0], colors=sl[1], count=dummy, sync=False)
            local[it['pos']] = sl
        if dest is not None:
            bdist.peer_fence(dev)                             # the back-projection kernels WERE the transfer
            how = 'peer stores into rank 0 (symmetric memory over NVLink)'
        else:
            bdist.gather_into(all_counts, local, buf_p if rank == 0 else None, buf_c if rank == 0 else None)
            how = 'NCCL point-to-point into the final slices' if ws > 1 else 'single rank'
        if rank != 0:
            return None, None, all_counts, how
        return buf_p, buf_c, all_counts, how

    def mask_backproject(self, depth: torch.Tensor, conf: torch.Tensor, bgr: torch.Tensor, Kinv64: np.ndarray,
                         R: np.ndarray, t: np.ndarray, min_conf: float, mode: int,
                         points: Optional[torch.Tensor] = None, colors: Optional[torch.Tensor] = None,
                         sync: bool = True, slot: int = 0, count: Optional[torch.Tensor] = None):
        """Mask + ordered compaction + back-projection (mvs_patchmatch.py:545-565 / dense_stereo.py:414-437).
        Returns (points (M,3) f64, colors (M,3) u8) device tensors; synchronises to read M.
        With sync=False nothing is read back: returns (points buffer, colors buffer, count tensor)."""
        H, W = depth.shape
        if points is None:
            points = torch.empty((H * W, 3), dtype=torch.float64, device=self.device)
        if colors is None:
            colors = torch.empty((H * W, 3), dtype=torch.uint8, device=self.device)
        if count is None:
            count = torch.zeros(1, dtype=torch.int64, device=self.device)
        Ki = np.ascontiguousarray(Kinv64, dtype=np.float64)
        Rm = np.ascontiguousarray(R, dtype=np.float64)
        tv = np.ascontiguousarray(t, dtype=np.float64).ravel()
        dp = C.POINTER(C.c_double)
        _cabi.check(self.lib.b200d_mask_backproject(_ptr(depth), _ptr(conf), _ptr(bgr), H, W, Ki.ctypes.data_as(dp),
                                                    Rm.ctypes.data_as(dp), tv.ctypes.data_as(dp), float(min_conf), mode,
                                                    _ptr(points), _ptr(colors), _ptr(count),
                                                    _ptr(self.scratch('fuse', H, W, slot)), self._stream()))
        self.launches += 3
        if not sync:
            return points, colors, count
        m = int(count.item())
        return points[:m], colors[:m]
