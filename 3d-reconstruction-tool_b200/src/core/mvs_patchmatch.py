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
    reference's CPU run under the same torch.manual_seed), `seed`, `device`, and two opt-in extensions the
    reference has no counterpart for: `consistency='geometric'` and `schedule='redblack'` (checkerboard propagation).
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
                 geo_rel_thresh: float = 0.01, schedule: str = 'reference'):
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
        if schedule not in ('reference', 'redblack'):
            raise ValueError("schedule must be 'reference' (the reference's Jacobi shifts) or 'redblack' (extension)")
        self.schedule = schedule
        self.geo_min_views, self.geo_px_thresh, self.geo_rel_thresh = geo_min_views, geo_px_thresh, geo_rel_thresh
        self.device = torch.device(device) if device is not None else torch.device('cuda')
        self.cache_images = True            # keep uploaded grays on the device between views
        self.views_in_flight = 4            # independent views enqueued concurrently (streams)
        self.device_postfilter = True       # run _filter_points on the GPU (same selection, b200dense/postfilter.py)
        self.device_prepare = True          # run _prepare_images on the GPU (bit-identical to the OpenCV calls)
        self._engine = None
        self._dev_gray: Dict[int, Tuple[int, torch.Tensor]] = {}
        self._dev_maps: Dict[int, dict] = {}
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

    # ------------------------------------------------------------------ public API
    def reconstruct(self, images: List[dict], poses: Dict[int, CameraPose],
                    sparse_points: np.ndarray = None) -> Tuple[np.ndarray, np.ndarray]:
        """Depth maps for every posed view, fused into one cloud (reference :72-139).

        Views are independent, so they are (a) sharded round-robin over the ranks of an initialised
        torch.distributed job (one process per GPU; without one this is a single-GPU run) and (b) kept
        `views_in_flight` at a time on separate streams of each GPU.  Rank 0 returns the fused and filtered
        cloud in the reference's order (sorted views, row-major pixels); other ranks return (0,3) arrays."""
        from b200dense import dist as bdist
        rank, ws = bdist.world()
        print("\n" + "=" * 60)
        print("PATCHMATCH MULTI-VIEW STEREO")
        print(f"  Scale: {self.scale}x, Patch: {self.patch_size}, Iters: {self.num_iterations}")
        print("=" * 60)
        t0 = time.time()
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
        self._dev_maps.clear()
        eng = self.engine
        Kinv = np.linalg.inv(self.K_scaled)
        jobs, job_views = [], []
        for pos in range(rank, n_cams, ws):
            ref_idx = cam_indices[pos]
            src_indices = self._select_source_views(ref_idx, cam_indices, poses, k=self.num_src_views)
            if len(src_indices) < 2:
                print(f"  [{pos+1}/{n_cams}] Cam {ref_idx}: skipped (not enough neighbors)")
                continue
            jobs.append(lambda slot, r=ref_idx, s=src_indices: self._patchmatch_device(r, s, proc_images, poses, slot=slot))
            job_views.append((pos, ref_idx, src_indices))
        # host-drawn random numbers must be consumed in view order -> one view at a time in that mode
        results = eng.run_views(jobs, 1 if self.rng == 'torch' else self.views_in_flight)
        maps = {ref_idx: out for (_, ref_idx, _), out in zip(job_views, results)}

        if self.consistency == 'geometric':
            # EXTENSION (no reference counterpart): cross-view geometric check.  The depth maps of the other ranks'
            # views arrive by one NCCL all-gather; the check itself is again view-local.
            H, W = next(iter(proc_images.values()))['shape']
            all_depth = bdist.all_gather_maps(cam_indices, {v: o['depth'] for v, o in maps.items()}, (H, W), eng.device)
            computed = bdist.all_gather_maps(cam_indices, {v: torch.ones(1, device=eng.device) for v in maps}, (1,),
                                             eng.device)
            for _, ref_idx, src_indices in job_views:
                srcs = [i for i in src_indices if float(computed[i].item()) > 0]
                if not srcs:
                    maps[ref_idx]['confidence'].zero_()
                    continue
                from b200dense.engine import kinv32, make_cams
                cams = make_cams(self.K_scaled, kinv32(self.K_scaled), poses[ref_idx].R, poses[ref_idx].t,
                                 [poses[i].R for i in srcs], [poses[i].t for i in srcs])
                geo = eng.geo_consistency(maps[ref_idx]['depth'], [all_depth[i].contiguous() for i in srcs], cams,
                                          self.geo_px_thresh, self.geo_rel_thresh)
                conf = maps[ref_idx]['confidence']
                maps[ref_idx]['confidence'] = torch.where(geo >= float(self.geo_min_views), conf, torch.zeros_like(conf))

        local = {}
        for pos, ref_idx, _ in job_views:
            out = maps[ref_idx]
            bgr = self._color_on_device(proc_images, ref_idx)
            pts, cols = eng.mask_backproject(out['depth'], out['confidence'], bgr, Kinv, poses[ref_idx].R,
                                             poses[ref_idx].t, float(self.min_views), 0)
            local[ref_idx] = (pts.clone(), cols.clone())
            print(f"  [{pos+1}/{n_cams}] Cam {ref_idx}: {pts.shape[0]:,} valid pixels")

        print("\nFusing depth maps...")
        pts_d, col_d = bdist.gather_point_clouds(cam_indices, local, eng.device)
        if rank != 0:
            return np.array([]).reshape(0, 3), np.array([]).reshape(0, 3)
        print(f"  Raw points: {pts_d.shape[0]:,}")
        if pts_d.shape[0] == 0:
            points, colors = np.array([]).reshape(0, 3), np.array([]).reshape(0, 3)
        elif self.device_postfilter:
            from b200dense.postfilter import filter_points
            fp, fc = filter_points(pts_d, col_d)                 # same selection and order as _filter_points
            points, colors = fp.cpu().numpy(), fc.cpu().numpy()
            print(f"  After filtering: {len(points):,}")
        else:
            points, colors = self._filter_points(pts_d.cpu().numpy(), col_d.cpu().numpy())
            print(f"  After filtering: {len(points):,}")
        print(f"\nPatchMatch MVS completed in {time.time() - t0:.1f}s")
        return points, colors

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
        index order (stable sort) -- reference :193-223."""
        c0 = poses[ref_idx].center
        d0 = poses[ref_idx].R[2, :]
        scored = []
        for idx in all_indices:
            if idx == ref_idx:
                continue
            baseline = np.linalg.norm(poses[idx].center - c0)
            angle = np.degrees(np.arccos(np.clip(np.dot(d0, poses[idx].R[2, :]), -1, 1)))
            scored.append((idx, baseline * (1 - abs(angle - 20) / 60) if 5 < angle < 60 else 0))
        scored.sort(key=lambda p: p[1], reverse=True)
        return [idx for idx, _ in scored[:k]]

    # ------------------------------------------------------------------ device path
    def _camera_block(self, ref_idx, src_indices, poses):
        from b200dense.engine import kinv32, make_cams
        return make_cams(self.K_scaled, kinv32(self.K_scaled), poses[ref_idx].R, poses[ref_idx].t,
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
                           seed=self.seed, view_id=int(ref_idx), slot=slot,
                           schedule=1 if self.schedule == 'redblack' else 0, **kw)

    def compute_depth_maps(self, images: Dict, poses: Dict[int, CameraPose], ref_indices: List[int] = None,
                           sources: Dict[int, List[int]] = None) -> Dict[int, DepthNormalMap]:
        """Batch form of `_patchmatch_cuda`: host grays in, host maps out, for many reference views at once.

        `images` is the dict `_prepare_images` returns.  Every distinct gray is uploaded once per call, the views
        run `views_in_flight` at a time on separate streams, and each view's depth / normal / confidence come back
        through pinned staging on its own stream, so copies overlap the other views' kernels.  This is what
        `reconstruct()` does internally; views with fewer than two sources are skipped like there (:110-112)."""
        from b200dense.engine import PINNED
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
            # uploads go to a copy stream in order of first use; a view waits only for its own images, so the first
            # views start while the later views' images are still crossing PCIe
            ready = {}
            copy_stream = eng.copy_stream()
            copy_stream.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(copy_stream):
                for r, src in plan:
                    for i in [r] + src:
                        if i not in ready:
                            self._gray_on_device(images, i)
                            ready[i] = torch.cuda.Event()
                            ready[i].record(copy_stream)
            jobs = []
            for r, src in plan:
                def job(slot, r=r, src=src):
                    st = torch.cuda.current_stream(eng.device)
                    for i in [r] + src:
                        st.wait_event(ready[i])
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
            self._dev_maps[r] = dict(owner=id(dm.depth), depth=out['depth'], confidence=out['confidence'])
            maps[r] = dm
        return maps

    def _patchmatch_cuda(self, ref_idx: int, src_indices: List[int], images: Dict,
                         poses: Dict[int, CameraPose]) -> DepthNormalMap:
        """Same seam as the reference's `_patchmatch_cuda` (:225-321): host grays in, host maps out."""
        out = self._patchmatch_device(ref_idx, src_indices, images, poses)
        depth, normal, conf = self.engine.to_host([out['depth'], out['normal'], out['confidence']])
        dm = DepthNormalMap(depth=depth, normal=normal, confidence=conf)
        self._dev_maps[ref_idx] = dict(owner=id(dm.depth), depth=out['depth'], confidence=out['confidence'])
        return dm

    def _fuse_depth_maps(self, depth_maps: Dict[int, DepthNormalMap], images: Dict,
                         poses: Dict[int, CameraPose]) -> Tuple[np.ndarray, np.ndarray]:
        """Mask conf >= min_views, back-project, concatenate in dict order (reference :536-570)."""
        eng = self.engine
        Kinv = np.linalg.inv(self.K_scaled)
        all_points, all_colors = [], []
        for idx, dm in depth_maps.items():
            dev = self._dev_maps.get(idx)
            if dev is not None and dev['owner'] == id(dm.depth):
                depth, conf = dev['depth'], dev['confidence']
            else:
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
