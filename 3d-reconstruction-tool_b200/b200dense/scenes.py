"""Synthetic textured scenes with analytically known poses.

The reference ships no images, poses or depth maps, so every
test, golden fixture and benchmark in this repo is fed by this generator.  It
produces exactly what the reference's dense stage consumes
(`/root/reference/src/core/sfm_pipeline.py:114-120`, `run_reconstruction.py:109-116`):

* ``images``: list of ``{'image': BGR uint8 (H,W,3), 'gray': uint8 (H,W)}``
* ``poses``:  ``{idx: (R (3,3) f64, t (3,) f64)}`` world->camera, ``X_cam = R X_w + t``
* ``K``:      3x3 f64 pinhole intrinsics, zero distortion

Rendering is exact ray casting (ray/plane or ray/sphere) followed by a bilinear
lookup into a band-limited random colour texture.  It is written with torch ops
so that large benchmark scenes can be synthesised on the GPU; this is data
synthesis, not part of the measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class Scene:
    kind: str
    K: np.ndarray                              # (3,3) f64
    images: List[dict]                         # [{'image': BGR u8, 'gray': u8}]
    poses: Dict[int, Tuple[np.ndarray, np.ndarray]]
    depth_gt: List[np.ndarray] = field(default_factory=list)   # per view (H,W) f32, z in camera
    depth_range: Tuple[float, float] = (0.1, 100.0)

    @property
    def shape(self) -> Tuple[int, int]:
        return next(im for im in self.images if im is not None)['image'].shape[:2]


def make_texture(seed: int = 0, size: int = 2048, sigma: float = 1.5) -> torch.Tensor:
    """(size,size,3) f32 in [0,255]: uniform noise, Gaussian-blurred, min-max stretched."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    tex = torch.rand(3, 1, size, size, generator=g)
    r = int(math.ceil(3 * sigma))
    xs = torch.arange(-r, r + 1, dtype=torch.float32)
    w = torch.exp(-0.5 * (xs / sigma) ** 2)
    w = w / w.sum()
    tex = F.conv2d(F.pad(tex, (r, r, 0, 0), mode='circular'), w.view(1, 1, 1, -1))
    tex = F.conv2d(F.pad(tex, (0, 0, r, r), mode='circular'), w.view(1, 1, -1, 1))
    tex = tex[:, 0]
    lo = tex.amin(dim=(1, 2), keepdim=True)
    hi = tex.amax(dim=(1, 2), keepdim=True)
    tex = (tex - lo) / (hi - lo) * 255.0
    return tex.permute(1, 2, 0).contiguous()


def _bilinear_wrap(tex: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Sample (S,S,3) texture at float texel coords with wrap-around."""
    S = tex.shape[0]
    u0 = torch.floor(u)
    v0 = torch.floor(v)
    fu = (u - u0).unsqueeze(-1)
    fv = (v - v0).unsqueeze(-1)
    x0 = u0.long() % S
    y0 = v0.long() % S
    x1 = (x0 + 1) % S
    y1 = (y0 + 1) % S
    a = tex[y0, x0] * (1 - fu) + tex[y0, x1] * fu
    b = tex[y1, x0] * (1 - fu) + tex[y1, x1] * fu
    return a * (1 - fv) + b * fv


def look_at_pose(center: np.ndarray, target: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """World->camera (R,t) with +z towards ``target`` and image y pointing down (world +y)."""
    z = target - center
    z = z / np.linalg.norm(z)
    up = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=0)
    t = -R @ center
    return R, t


def _rays(K: np.ndarray, R: np.ndarray, H: int, W: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """World-space ray directions with unit camera-z, per pixel (H,W,3) f64, plus cam-frame dirs."""
    ys, xs = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float64),
                            torch.arange(W, device=device, dtype=torch.float64), indexing='ij')
    dc = torch.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], torch.ones_like(xs)], dim=-1)
    Rt = torch.from_numpy(np.ascontiguousarray(R.T)).to(device)
    dw = dc @ Rt.T          # X_w = R^T X_c
    return dw, dc


def make_scene(kind: str = 'plane', n_views: int = 5, width: int = 640, height: int = 480,
               focal: float = None, arc_deg: float = 8.0, seed: int = 0,
               device: str = 'cpu', radius: float = 5.0, render=None) -> Scene:
    """Build a synthetic multi-view scene.

    ``plane``     : textured plane z=5, cameras on an arc of ``arc_deg`` per step around it.
    ``wavy``      : height field z = 5 + 0.25 sin(1.3X) cos(1.1Y), same cameras.
    ``turntable`` : textured unit-ish sphere (r=1.5) inside a large textured shell (r=12),
                    cameras on a full circle of ``radius`` looking at the origin,
                    ``arc_deg`` between neighbouring views.
    ``render``: optional iterable of view indices to rasterise; the other views get poses only
    (their ``images`` / ``depth_gt`` entries are None) -- large benchmark scenes render what a rank needs.
    """
    W, H = int(width), int(height)
    f = float(focal) if focal is not None else 0.9 * W
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1.0]], dtype=np.float64)
    dev = torch.device(device)
    tex = make_texture(seed).to(dev).double()
    # ~0.6 texels per pixel at the nominal scene depth: image features are a few pixels wide
    tex_scale = 0.6 * f / radius

    poses: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}
    images: List[dict] = []
    depth_gt: List[np.ndarray] = []

    for i in range(n_views):
        ang = math.radians((i - (n_views - 1) / 2.0) * arc_deg)
        if kind in ('plane', 'wavy'):
            target = np.array([0.0, 0.0, radius])
            C = np.array([radius * math.sin(ang), 0.0, radius - radius * math.cos(ang)])
        elif kind == 'turntable':
            target = np.zeros(3)
            C = np.array([radius * math.sin(ang), -0.6, -radius * math.cos(ang)])
        else:
            raise ValueError(f"unknown scene kind {kind!r}")
        R, t = look_at_pose(C, target)
        poses[i] = (R, t)
        if render is not None and i not in render:
            images.append(None)
            depth_gt.append(None)
            continue

        dw, _ = _rays(K, R, H, W, dev)
        Ct = torch.from_numpy(C).to(dev)
        if kind == 'plane':
            s = (radius - Ct[2]) / dw[..., 2]
            P = Ct + s.unsqueeze(-1) * dw
            u = P[..., 0] * tex_scale + 1024.0
            v = P[..., 1] * tex_scale + 1024.0
        elif kind == 'wavy':
            s = (radius - Ct[2]) / dw[..., 2]
            for _ in range(25):     # fixed-point ray / height-field intersection
                P = Ct + s.unsqueeze(-1) * dw
                zt = radius + 0.25 * torch.sin(1.3 * P[..., 0]) * torch.cos(1.1 * P[..., 1])
                s = s + (zt - P[..., 2]) / dw[..., 2]
            P = Ct + s.unsqueeze(-1) * dw
            u = P[..., 0] * tex_scale + 1024.0
            v = P[..., 1] * tex_scale + 1024.0
        else:
            # nearest positive hit with the inner sphere, else the enclosing shell
            a = (dw * dw).sum(-1)
            b = 2.0 * (dw * Ct).sum(-1)
            cc_in = (Ct * Ct).sum() - 1.5 ** 2
            cc_out = (Ct * Ct).sum() - 12.0 ** 2
            disc_in = b * b - 4 * a * cc_in
            s_in = (-b - torch.sqrt(disc_in.clamp_min(0))) / (2 * a)
            hit = (disc_in > 0) & (s_in > 0)
            s_out = (-b + torch.sqrt(b * b - 4 * a * cc_out)) / (2 * a)
            s = torch.where(hit, s_in, s_out)
            P = Ct + s.unsqueeze(-1) * dw
            lon = torch.atan2(P[..., 0], -P[..., 2])
            lat = torch.asin((P[..., 1] / P.norm(dim=-1)).clamp(-1, 1))
            scale = torch.where(hit, torch.tensor(0.257 * f, device=dev, dtype=torch.float64),
                                torch.tensor(0.6 * f, device=dev, dtype=torch.float64))
            u = lon * scale + 1024.0
            v = lat * scale + torch.where(hit, 1024.0, 333.0)
        col = _bilinear_wrap(tex, u, v)
        bgr = col.round().clamp(0, 255).to(torch.uint8)
        gray = (0.114 * col[..., 0] + 0.587 * col[..., 1] + 0.299 * col[..., 2]).round().clamp(0, 255).to(torch.uint8)
        images.append({'path': f'synthetic_{i:04d}', 'image': bgr.cpu().numpy(), 'gray': gray.cpu().numpy()})
        depth_gt.append(s.float().cpu().numpy())    # dw has unit camera-z, so s is the camera depth

    if kind == 'turntable':
        rng = (max(0.5, radius - 2.5), radius + 13.0)
    else:
        rng = (0.4 * radius, 2.0 * radius)
    return Scene(kind=kind, K=K, images=images, poses=poses, depth_gt=depth_gt, depth_range=rng)


def gray_f32(scene: Scene) -> List[np.ndarray]:
    """Working-resolution grayscale in [0,1] f32, computed like the reference's
    `_prepare_images` at scale 1.0 (`mvs_patchmatch.py:176-177`): OpenCV BGR2GRAY then /255."""
    import cv2 as cv
    return [None if im is None else cv.cvtColor(im['image'], cv.COLOR_BGR2GRAY).astype(np.float32) / 255.0
            for im in scene.images]
