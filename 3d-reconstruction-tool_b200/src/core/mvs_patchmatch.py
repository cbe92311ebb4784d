"""PatchMatch multi-view stereo -- drop-in for the reference's `core.mvs_patchmatch`
(reference: src/core/mvs_patchmatch.py), computed by hand-written sm_100a kernels.

Same class names, constructor keywords, method names and return types as the reference
(`PatchMatchMVS`, `DepthNormalMap`; `reconstruct`, `_patchmatch_cuda`, `_fuse_depth_maps`, ...), same
stdout banners (the reference's GUI greps them, gui/worker.py:112-122), same error conventions
(fewer than 3 cameras -> two empty arrays, never an exception for data-dependent emptiness).

What differs, by design:
  * every per-pixel operation runs in libb200dense.so (include/b200dense.h); there is no CPU
    fallback -- without a CUDA device the compute methods raise B200DenseError;
  * the whole `_patchmatch_cuda` body is 33 kernel launches (the reference issues ~7-8k);
  * extra keyword-only constructor arguments: `num_src_views` (the reference hard-codes k=4 at
    :108), `rng` ('philox' = counter-based in-kernel generator; 'torch' = draw on the host with
    torch.rand/randn in the reference's call order, which makes results bit-identical to the
    reference's CPU run under the same torch.manual_seed), `seed`, `device`, `arith` ('parity' = bit-exact,
    the default; 'fast' = the opt-in re-associated arithmetic, NOT bit-exact, see include/b200dense.h), and
    three opt-in extensions the reference has no counterpart for: `consistency='geometric'` (cross-view
    depth-map check), `schedule='redblack'` (checkerboard propagation) and `schedule='viewprop'` (after every
    iteration the source views' current hypotheses are proposed to the reference view; the depth and normal
    maps of all views are exchanged by an NCCL all-gather when the views are sharded over GPUs).
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import cv2 as cv
import numpy as np
import torch

from .camera import Camera, CameraPose


@dataclass
class DepthNormalMap:
    """Depth (H,W), normal (H,W,3) and confidence (H,W) of one view (reference :30-35)."""
    depth: np.ndarray
    normal: np.ndarray
    confidence: np.ndarray


class PatchMatchMVS:
    def __init__(self, camera: Camera, scale: float = 0.25, patch_size: int = 11, num_iterations: int = 3,
                 num_samples: int = 8, min_views: int = 3, depth_min: float = 0.1, depth_max: float = 100.0,
                 *, num_src_views: int = 4, rng: str = 'philox', seed: int = 0, device=None,
                 consistency: str = 'photometric', geo_min_views: int = 2, geo_px_thresh: float = 1.0,
                 geo_rel_thresh: float = 0.01, schedule: str = 'reference', arith: str = 'parity'):
        self.camera = camera
        self.scale = scale
        self.patch_size = patch_size
        self.num_iterations = num_iterations
        self.num_samples = num_samples
        self.min_views = min_views
        self.depth_min = depth_min
        self.depth_max = depth_max
        self.num_src_views = num_src_views
        if rng not in ('philox', 'torch'):
            raise ValueError("rng must be 'philox' or 'torch'")
        self.rng = rng
        self.seed = seed
        if consistency not in ('photometric', 'geometric'):
            raise ValueError("consistency must be 'photometric' (the reference's check) or 'geometric' (extension)")
        self.consistency = consistency
        if schedule not in ('reference', 'redblack', 'viewprop'):
            raise ValueError("schedule must be 'reference' (the reference's Jacobi shifts), 'redblack' or 'viewprop' "
                             "(extensions)")
        self.schedule = schedule
        if arith not in ('parity', 'fast'):
            raise ValueError("arith must be 'parity' (bit-exact, default) or 'fast' (opt-in, not bit-exact)")
        self.arith = arith
        self.geo_min_views, self.geo_px_thresh, self.geo_rel_thresh = geo_min_views, geo_px_thresh, geo_rel_thresh
        self.device = torch.device(device) if device is not None else torch.device('cuda')
        self.cache_images = True            # keep uploaded grays on the device between views
        self.views_in_flight = 4            # independent views enqueued concurrently (streams)
        self.device_postfilter = True       # run _filter_points on the GPU (same selection, b200dense/postfilter.py)
        self.device_prepare = True          # run _prepare_images on the GPU (bit-identical to the OpenCV calls)
        self._engine = None
        self._dev_gray: Dict[int, Tuple[int, torch.Tensor]] = {}
        self._kinv32 = None
        self.last_timings: Dict[str, float] = {}   # phase breakdown of the last reconstruct() (seconds)
        if torch.cuda.is_available():
            print(f"PatchMatch MVS using GPU: {torch.cuda.get_device_name()}")
        else:
            print("PatchMatch MVS: no CUDA device visible (compute calls will raise; there is no CPU path)")
        # intrinsics at working resolution (reference :68-70 scales the first two rows)
        self.K_scaled = camera.K.copy()
        self.K_scaled[:2] *= scale

    # ------------------------------------------------------------------ plumbing
    @property
    def engine(self):
        if self._engine is None:
            from b200dense.engine import DenseEngine
            self._engine = DenseEngine(self.device)
        return self._engine

    def _prepare_images_device(self, images: List[dict], indices: List[int]) -> Dict:
        """Device form of `_prepare_images`: entries are filled lazily (first use uploads the photograph as uint8 and
        runs the bit-exact resize + gray kernel), so a rank only prepares the views it touches."""
        out = {}
        for idx in indices:
            h, w = images[idx]['image'].shape[:2]
            out[idx] = {'image': images[idx]['image'], 'shape': (int(h * self.scale), int(w * self.scale))}
        return out

    def _ensure_device_image(self, entry: dict) -> None:
        if 'gray_dev' not in entry:
            eng = self.engine
            nh, nw = entry['shape']
            entry['color_dev'], entry['gray_dev'] = eng.prepare_image(eng.to_device(entry['image'], torch.uint8), nh, nw)

    def _color_on_device(self, images: Dict, idx: int) -> torch.Tensor:
        entry = images[idx]
        if 'image' in entry and 'color' not in entry:
            self._ensure_device_image(entry)
            return entry['color_dev']
        return self.engine.to_device(entry['color'], torch.uint8)

    def _gray_on_device(self, images: Dict, idx: int) -> torch.Tensor:
        if 'gray' not in images[idx]:
            self._ensure_device_image(images[idx])
            return images[idx]['gray_dev']
        g = images[idx]['gray']
        hit = self._dev_gray.get(idx)
        if self.cache_images and hit is not None and hit[0] == id(g):
            return hit[1]
        t = self.engine.to_device(g)
        if self.cache_images:
            self._dev_gray[idx] = (id(g), t)
        return t

    def _stage_images(self, images: Dict, indices, ready: Dict = None) -> Dict[int, "torch.cuda.Event"]:
        """Upload (and, on the device-prepare path, resize + gray) the listed images on the copy stream, in order of
        first use, and return one event per image.  A job on another stream waits for the events of the images it
        reads before it launches anything, so no kernel can read an image whose upload / prepare kernel is still in
        flight on a different stream.  Called incrementally (`ready` carries over): reconstruct() stages the images of
        view j right before it enqueues view j, so the GPU already works on the first views while the host is still
        pushing later photographs through PCIe (uploads from the caller's pageable arrays block the host thread)."""
        eng = self.engine
        copy_stream = eng.copy_stream()
        if ready is None:
            ready = {}
            copy_stream.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(copy_stream):
            for i in indices:
                if i in ready:
                    continue
                if 'gray' in images[i] and not self.cache_images:
                    continue                             # uncached host grays are uploaded by the job, on its own stream
                self._gray_on_device(images, i)
                ready[i] = torch.cuda.Event()
                ready[i].record(copy_stream)
        return ready

    @staticmethod
    def _wait_images(ready: Dict, indices) -> None:
        st = torch.cuda.current_stream()
        for i in indices:
            ev = ready.get(i)
            if ev is not None:
                st.wait_event(ev)

    # ------------------------------------------------------------------ public API
    def reconstruct(self, images: List[dict], poses: Dict[int, CameraPose],
                    sparse_points: np.ndarray = None) -> Tuple[np.ndarray, np.ndarray]:
        """Depth maps for every posed view, fused into one cloud (reference :72-139).

        Views are independent, so they are (a) sharded in contiguous blocks over the ranks of an initialised
        torch.distributed job (one process per GPU; without one this is a single-GPU run) and (b) kept
        `views_in_flight` at a time on separate streams of each GPU.  Rank 0 returns the fused and filtered
        cloud in the reference's order (sorted views, row-major pixels); other ranks return (0,3) arrays.
        The fusion stage synchronises with the host once (to size the cloud): every view's mask is counted on the
        device, then each view is back-projected straight into its slice of the final buffer (DenseEngine.fuse_to_rank0)
        -- in a multi-GPU job the other ranks' kernels store into rank 0's buffer through NVLink peer mappings (NCCL
        point-to-point into the final slices where peer memory is unavailable)."""
        from b200dense import dist as bdist
        rank, ws = bdist.world()
        print("\n" + "=" * 60)
        print("PATCHMATCH MULTI-VIEW STEREO")
        print(f"  Scale: {self.scale}x, Patch: {self.patch_size}, Iters: {self.num_iterations}")
        print("=" * 60)
        t0 = time.time()
        tm = self.last_timings = {}
        cam_indices = sorted(poses.keys())
        n_cams = len(cam_indices)
        if n_cams < 3:
            print("Need at least 3 cameras")
            return np.array([]), np.array([])

        self._estimate_depth_range(poses, sparse_points)
        print(f"  Depth range: [{self.depth_min:.2f}, {self.depth_max:.2f}]")
        print("\nPreparing images...")
        proc_images = (self._prepare_images_device(images, cam_indices) if self.device_prepare
                       else self._prepare_images(images, cam_indices))

        print(f"\nComputing depth maps for {n_cams} views...")
        self._dev_gray.clear()
        eng = self.engine
        dev = eng.device
        Kinv = np.linalg.inv(self.K_scaled)
        ready = self._stage_images(proc_images, [])             # opens the copy stream; images follow view by view
        # host-drawn random numbers must be consumed in view order -> one view at a time in that mode
        in_flight = 1 if self.rng == 'torch' else self.views_in_flight
        mine = [(pos, cam_indices[pos]) for pos in bdist.shard_positions(n_cams, rank, ws)]
        if self.schedule == 'viewprop':
            plan = []
            for pos, ref_idx in mine:
                src_indices = self._select_source_views(ref_idx, cam_indices, poses, k=self.num_src_views)
                if len(src_indices) < 2:
                    print(f"  [{pos+1}/{n_cams}] Cam {ref_idx}: skipped (not enough neighbors)")
                    continue
                plan.append((pos, ref_idx, src_indices))
            maps = self._run_viewprop(plan, cam_indices, proc_images, poses, ready, in_flight)
        else:
            # source selection, image staging and kernel launches of view j+1 all happen on the host while the GPU
            # works on view j: nothing but the first view's preparation is exposed
            def view_job(pos, ref_idx):
                def job(slot):
                    h0 = time.perf_counter()
                    src_indices = self._select_source_views(ref_idx, cam_indices, poses, k=self.num_src_views)
                    if len(src_indices) < 2:
                        print(f"  [{pos+1}/{n_cams}] Cam {ref_idx}: skipped (not enough neighbors)")
                        return None
                    h1 = time.perf_counter()
                    self._stage_images(proc_images, [ref_idx] + list(src_indices), ready)
                    self._wait_images(ready, [ref_idx] + list(src_indices))
                    h2 = time.perf_counter()
                    out = self._patchmatch_device(ref_idx, src_indices, proc_images, poses, slot=slot)
                    h3 = time.perf_counter()
                    tm['host_select_s'] = tm.get('host_select_s', 0.0) + (h1 - h0)
                    tm['host_stage_upload_s'] = tm.get('host_stage_upload_s', 0.0) + (h2 - h1)
                    tm['host_enqueue_s'] = tm.get('host_enqueue_s', 0.0) + (h3 - h2)
                    # the fusion needs depth + confidence only; at 3840x2160 a normal map is 100 MB per view, which a
                    # 200-view job would otherwise hold until the end (released in stream order: allocator-safe)
                    out.pop('normal', None)
                    return src_indices, out
                return job
            results = eng.run_views([view_job(pos, r) for pos, r in mine], in_flight)
            plan = [(pos, r, res[0]) for (pos, r), res in zip(mine, results) if res is not None]
            maps = {r: res[1] for (_, r), res in zip(mine, results) if res is not None}
        if self.consistency == 'geometric':
            self._geometric_filter(plan, cam_indices, proc_images, poses, maps)

        # ---- fusion: count every mask on the device, read the counts once, back-project into place (engine.fuse_to_rank0)
        items = [dict(pos=pos, depth=maps[r]['depth'], conf=maps[r]['confidence'], R=poses[r].R, t=poses[r].t,
                      bgr=(lambda r=r: self._color_on_device(proc_images, r))) for pos, r, _ in plan]
        torch.cuda.current_stream(dev).synchronize()           # phase boundary for the timing breakdown
        t1 = time.time()
        tm['depth_maps_s'] = t1 - t0
        buf_p, buf_c, all_counts, how = eng.fuse_to_rank0(n_cams, items, Kinv, float(self.min_views), 0)
        for pos, ref_idx, _ in plan:
            print(f"  [{pos+1}/{n_cams}] Cam {ref_idx}: {int(all_counts[pos]):,} valid pixels")
        print("\nFusing depth maps...")
        tm['gather'] = how
        tm['fusion_gather_s'] = time.time() - t1
        if rank != 0:
            torch.cuda.current_stream(dev).synchronize()
            return np.array([]).reshape(0, 3), np.array([]).reshape(0, 3)
        pts_d, col_d = buf_p, buf_c
        print(f"  Raw points: {pts_d.shape[0]:,}")
        t2 = time.time()
        if pts_d.shape[0] == 0:
            points, colors = np.array([]).reshape(0, 3), np.array([]).reshape(0, 3)
        elif self.device_postfilter:
            from b200dense.postfilter import filter_points
            fp, fc = filter_points(pts_d, col_d)                 # same selection and order as _filter_points
            # one-off size: a pageable download beats page-locking a fresh staging buffer (~1 GB/s to pin)
            points, colors = fp.cpu().numpy(), fc.cpu().numpy()
            print(f"  After filtering: {len(points):,}")
        else:
            points, colors = self._filter_points(pts_d.cpu().numpy(), col_d.cpu().numpy())
            print(f"  After filtering: {len(points):,}")
        tm['filter_download_s'] = time.time() - t2
        tm['raw_points'] = int(pts_d.shape[0])
        tm['total_s'] = time.time() - t0
        print(f"\nPatchMatch MVS completed in {time.time() - t0:.1f}s")
        return points, colors

    def _geometric_filter(self, plan, cam_indices, proc_images, poses, maps) -> None:
        """EXTENSION (no reference counterpart): cross-view geometric check.  The depth maps of the other ranks' views
        arrive by one NCCL all-gather; the check itself is again view-local.  Confidence is zeroed where fewer than
        `geo_min_views` sources agree geometrically."""
        from b200dense import dist as bdist
        eng = self.engine
        H, W = next(iter(proc_images.values()))['shape']
        all_depth = bdist.all_gather_maps(cam_indices, {v: o['depth'] for v, o in maps.items()}, (H, W), eng.device)
        computed = bdist.all_gather_maps(cam_indices, {v: torch.ones(1, device=eng.device) for v in maps}, (1,),
                                         eng.device)
        has_map = {v: float(t.item()) > 0 for v, t in computed.items()}
        for _, ref_idx, src_indices in plan:
            srcs = [i for i in src_indices if has_map[i]]
            if not srcs:
                maps[ref_idx]['confidence'].zero_()
                continue
            cams = self._camera_block(ref_idx, srcs, poses)
            geo = eng.geo_consistency(maps[ref_idx]['depth'], [all_depth[i].contiguous() for i in srcs], cams,
                                      self.geo_px_thresh, self.geo_rel_thresh)
            conf = maps[ref_idx]['confidence']
            maps[ref_idx]['confidence'] = torch.where(geo >= float(self.geo_min_views), conf, torch.zeros_like(conf))

    def _run_viewprop(self, plan, cam_indices, images, poses, ready, in_flight) -> Dict[int, dict]:
        """EXTENSION `schedule='viewprop'` (the reference's README names view propagation, its code has none): all
        views advance in lock step.  After every PatchMatch iteration the depth AND normal maps of all views are
        exchanged (NCCL all-gather when views are sharded; the snapshot makes the result independent of the number of
        GPUs) and each view evaluates, per source view, the hypothesis that source currently holds for the same
        surface point (b200d_pm_viewprop); then the next iteration.  The confidence pass runs at the end."""
        from b200dense import _cabi, dist as bdist
        eng = self.engine
        dev = eng.device
        st, rngs = {}, {}
        arith = self._arith_code()
        for _, r, _ in plan:
            H, W = images[r]['shape']
            st[r] = dict(depth=torch.empty((H, W), device=dev), normal=torch.empty((H, W, 3), device=dev),
                         confidence=torch.empty((H, W), device=dev), best_cost=torch.empty((H, W), device=dev))
            rngs[r] = ({k: v.to(dev, non_blocking=True).contiguous() for k, v in self._host_rng(H, W).items()}
                       if self.rng == 'torch' else {})
        H, W = next(iter(images.values()))['shape']
        iters = self.num_iterations

        def pm_stage(r, src, it_begin, it_end, stage):
            def job(slot):
                self._stage_images(images, [r] + list(src), ready)
                self._wait_images(ready, [r] + list(src))
                return eng.pm_view(self._gray_on_device(images, r), [self._gray_on_device(images, i) for i in src],
                                   self._camera_block(r, src, poses), self.patch_size, iters, self.num_samples,
                                   self.depth_min, self.depth_max, seed=self.seed, view_id=int(r), slot=slot, out=st[r],
                                   arith=arith, it_begin=it_begin, it_end=it_end, stage=stage, **rngs[r])
            return job

        if iters == 0:
            eng.run_views([pm_stage(r, src, 0, 0, 0) for _, r, src in plan], in_flight)
            return st
        for it in range(iters):
            stage = _cabi.STAGE_NO_CONF | (_cabi.STAGE_NO_INIT if it > 0 else 0)
            eng.run_views([pm_stage(r, src, it, it + 1, stage) for _, r, src in plan], in_flight)
            all_depth = bdist.all_gather_maps(cam_indices, {r: st[r]['depth'] for r in st}, (H, W), dev)
            all_normal = bdist.all_gather_maps(cam_indices, {r: st[r]['normal'] for r in st}, (H, W, 3), dev)

            def vp_job(r, src):
                def job(slot):
                    eng.pm_viewprop(self._gray_on_device(images, r), [self._gray_on_device(images, i) for i in src],
                                    self._camera_block(r, src, poses), self.patch_size, self.depth_min, self.depth_max,
                                    [all_depth[i] for i in src], [all_normal[i] for i in src], st[r], arith=arith,
                                    slot=slot)
                return job
            eng.run_views([vp_job(r, src) for _, r, src in plan], in_flight)
        eng.run_views([pm_stage(r, src, iters, 0, _cabi.STAGE_NO_INIT) for _, r, src in plan], in_flight)
        return st

    # ------------------------------------------------------------------ host heuristics
    def _estimate_depth_range(self, poses: Dict[int, CameraPose], sparse_points: np.ndarray = None):
        """Percentiles 1 / 99*1.5 of the sparse points' camera depths, else a multiple of the camera
        spread (reference :141-165).  Sets self.depth_min / self.depth_max."""
        if sparse_points is not None and len(sparse_points) > 0:
            zs = []
            for idx in poses:
                z = poses[idx].transform_points(sparse_points)[:, 2]
                front = z > 0
                if np.any(front):
                    zs.extend(z[front])
            if zs:
                self.depth_min = max(0.1, np.percentile(zs, 1))
                self.depth_max = np.percentile(zs, 99) * 1.5
                return
        centers = np.array([poses[i].center for i in poses])
        spread = np.percentile(np.linalg.norm(centers - np.median(centers, axis=0), axis=1), 90)
        self.depth_min = max(0.1, spread * 0.05)
        self.depth_max = spread * 10.0

    def _prepare_images(self, images: List[dict], indices: List[int]) -> Dict:
        """Working-resolution colour + [0,1] gray per view (reference :167-191; the Sobel gradients
        the reference also computes are never read and are not produced here)."""
        out = {}
        for idx in indices:
            img = images[idx]['image']
            h, w = img.shape[:2]
            nh, nw = int(h * self.scale), int(w * self.scale)
            small = cv.resize(img, (nw, nh))
            gray = cv.cvtColor(small, cv.COLOR_BGR2GRAY).astype(np.float32) / 255.0
            out[idx] = {'color': small, 'gray': gray, 'shape': (nh, nw)}
        return out

    def _select_source_views(self, ref_idx: int, all_indices: List[int], poses: Dict[int, CameraPose],
                             k: int = 4) -> List[int]:
        """Top-k by baseline * (1 - |angle - 20|/60) for 5 < angle < 60 degrees, else 0; ties keep
        index order (stable sort) -- reference :193-223.  Camera centres and viewing directions are computed once per
        pose set (the reference recomputes `-R.T @ t` for every pair; same values, 200 views = 40,000 pairs)."""
        cache = getattr(self, '_pose_cache', None)
        if cache is None or cache[0] is not poses or cache[1] != len(poses):
            cache = (poses, len(poses), {i: poses[i].center for i in poses}, {i: poses[i].R[2, :] for i in poses})
            self._pose_cache = cache
        centers, dirs = cache[2], cache[3]
        c0 = centers[ref_idx]
        d0 = dirs[ref_idx]
        scored = []
        for idx in all_indices:
            if idx == ref_idx:
                continue
            baseline = np.linalg.norm(centers[idx] - c0)
            angle = np.degrees(np.arccos(np.clip(np.dot(d0, dirs[idx]), -1, 1)))
            scored.append((idx, baseline * (1 - abs(angle - 20) / 60) if 5 < angle < 60 else 0))
        scored.sort(key=lambda p: p[1], reverse=True)
        return [idx for idx, _ in scored[:k]]

    # ------------------------------------------------------------------ device path
    def _arith_code(self) -> int:
        return 1 if self.arith == 'fast' else 0

    def _camera_block(self, ref_idx, src_indices, poses):
        from b200dense.engine import kinv32, make_cams
        key = self.K_scaled.tobytes()
        if self._kinv32 is None or self._kinv32[0] != key:      # one fp32 inverse per intrinsics, not per view
            self._kinv32 = (key, kinv32(self.K_scaled))
        return make_cams(self.K_scaled, self._kinv32[1], poses[ref_idx].R, poses[ref_idx].t,
                         [poses[i].R for i in src_indices], [poses[i].t for i in src_indices])

    def _host_rng(self, H: int, W: int) -> dict:
        """Draw on the host in the reference's call order (:271, :279-280, :471, :475) and build the
        initial hypotheses with the same torch ops (:268-281)."""
        import torch.nn.functional as F
        iu, nx, ny = torch.rand(H, W), torch.randn(H, W), torch.randn(H, W)
        n = self.num_iterations * self.num_samples
        ru = torch.empty(n, H, W)
        rn = torch.empty(n, H, W, 3)
        for j in range(n):
            ru[j] = torch.rand(H, W)
            rn[j] = torch.randn(H, W, 3)
        ld0, ld1 = np.log(self.depth_min), np.log(self.depth_max)
        depth0 = torch.exp(iu * (ld1 - ld0) + ld0)
        n0 = torch.zeros(H, W, 3)
        n0[:, :, 2] = -1
        n0[:, :, 0] = nx * 0.3
        n0[:, :, 1] = ny * 0.3
        return dict(init_depth=depth0, init_normal=F.normalize(n0, dim=-1), rand_u=ru, rand_n=rn)

    def _patchmatch_device(self, ref_idx: int, src_indices: List[int], images: Dict,
                           poses: Dict[int, CameraPose], slot: int = 0) -> dict:
        """PatchMatch of one view; returns DEVICE tensors dict(depth, normal, confidence)."""
        eng = self.engine
        ref = self._gray_on_device(images, ref_idx)
        srcs = [self._gray_on_device(images, i) for i in src_indices]
        H, W = images[ref_idx]['shape']
        kw = {}
        if self.rng == 'torch':
            kw = {k: v.to(eng.device, non_blocking=True).contiguous() for k, v in self._host_rng(H, W).items()}
        return eng.pm_view(ref, srcs, self._camera_block(ref_idx, src_indices, poses), self.patch_size,
                           self.num_iterations, self.num_samples, self.depth_min, self.depth_max,
                           seed=self.seed, view_id=int(ref_idx), slot=slot, arith=self._arith_code(),
                           schedule=1 if self.schedule == 'redblack' else 0, **kw)

    def compute_depth_maps(self, images: Dict, poses: Dict[int, CameraPose], ref_indices: List[int] = None,
                           sources: Dict[int, List[int]] = None) -> Dict[int, DepthNormalMap]:
        """Batch form of `_patchmatch_cuda`: host grays in, host maps out, for many reference views at once.

        `images` is the dict `_prepare_images` returns.  Every distinct gray is uploaded once per call, the views
        run `views_in_flight` at a time on separate streams, and each view's depth / normal / confidence come back
        through pinned staging on its own stream, so copies overlap the other views' kernels.  This is what
        `reconstruct()` does internally; views with fewer than two sources are skipped like there (:110-112)."""
        from b200dense.engine import PINNED
        if self.schedule == 'viewprop':
            raise ValueError("schedule='viewprop' couples the views: use reconstruct()")
        eng = self.engine
        all_indices = sorted(poses.keys())
        ref_indices = list(all_indices if ref_indices is None else ref_indices)
        plan = []
        for r in ref_indices:
            src = sources[r] if sources is not None else self._select_source_views(r, all_indices, poses,
                                                                                   k=self.num_src_views)
            if len(src) >= 2:
                plan.append((r, list(src)))
        keep = self.cache_images
        self.cache_images = True
        if not keep:
            self._dev_gray.clear()                       # re-upload: the caller asked for no caching across calls
        try:
            ready = self._stage_images(images, [i for r, src in plan for i in [r] + src])
            jobs = []
            for r, src in plan:
                def job(slot, r=r, src=src):
                    self._wait_images(ready, [r] + src)
                    out = self._patchmatch_device(r, src, images, poses, slot=slot)
                    host = []
                    for k in ('depth', 'normal', 'confidence'):
                        h = PINNED.take(out[k].shape, out[k].dtype)
                        h.copy_(out[k], non_blocking=True)
                        host.append(h)
                    return out, host
                jobs.append(job)
            results = eng.run_views(jobs, 1 if self.rng == 'torch' else self.views_in_flight)
            torch.cuda.current_stream().synchronize()
        finally:
            self.cache_images = keep
            if not keep:
                self._dev_gray.clear()
        maps = {}
        for (r, _), (out, host) in zip(plan, results):
            dm = DepthNormalMap(depth=PINNED.as_numpy(host[0]), normal=PINNED.as_numpy(host[1]),
                                confidence=PINNED.as_numpy(host[2]))
            maps[r] = dm
        return maps

    def _patchmatch_cuda(self, ref_idx: int, src_indices: List[int], images: Dict,
                         poses: Dict[int, CameraPose]) -> DepthNormalMap:
        """Same seam as the reference's `_patchmatch_cuda` (:225-321): host grays in, host maps out."""
        out = self._patchmatch_device(ref_idx, src_indices, images, poses)
        depth, normal, conf = self.engine.to_host([out['depth'], out['normal'], out['confidence']])
        dm = DepthNormalMap(depth=depth, normal=normal, confidence=conf)
        return dm

    def _fuse_depth_maps(self, depth_maps: Dict[int, DepthNormalMap], images: Dict,
                         poses: Dict[int, CameraPose]) -> Tuple[np.ndarray, np.ndarray]:
        """Mask conf >= min_views, back-project, concatenate in dict order (reference :536-570)."""
        eng = self.engine
        Kinv = np.linalg.inv(self.K_scaled)
        all_points, all_colors = [], []
        for idx, dm in depth_maps.items():
            # always from the caller's host arrays, like the reference: edits made to a map after it was
            # returned (own filtering, a replaced confidence) are honoured; the two maps are a few MB
            depth, conf = eng.to_device(dm.depth), eng.to_device(dm.confidence)
            bgr = eng.to_device(images[idx]['color'], torch.uint8)
            pts, cols = eng.mask_backproject(depth, conf, bgr, Kinv, poses[idx].R, poses[idx].t,
                                             float(self.min_views), 0)
            if pts.shape[0] == 0:
                continue
            all_points.append(pts.cpu().numpy())
            all_colors.append(cols.cpu().numpy())
        if not all_points:
            return np.array([]).reshape(0, 3), np.array([]).reshape(0, 3)
        return np.vstack(all_points), np.vstack(all_colors)

    def _filter_points(self, points: np.ndarray, colors: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Keep points closer to the median centroid than the 95th-percentile distance, then one point
        per 0.01 voxel: first occurrence, output ordered by voxel key (reference :572-588)."""
        dist = np.linalg.norm(points - np.median(points, axis=0), axis=1)
        keep = dist < np.percentile(dist, 95)
        points, colors = points[keep], colors[keep]
        v = np.floor(points / 0.01).astype(np.int64)
        keys = v[:, 0] * 1000000000 + v[:, 1] * 1000000 + v[:, 2]
        _, first = np.unique(keys, return_index=True)
        return points[first], colors[first]
