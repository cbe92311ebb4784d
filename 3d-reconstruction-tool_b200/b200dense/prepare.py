"""Device-side `_prepare_images`: resize + gray of the input photographs on the GPU,
bit-identical to the OpenCV calls the reference makes (mvs_patchmatch.py:176-177, dense_stereo.py:166-168).

`resize_tables` rebuilds, on the host, the per-column / per-row tables OpenCV's INTER_LINEAR resize builds for
uint8 images (source index and two 11-bit fixed-point coefficients); the kernel (csrc/prep_kernels.cuh) does the
integer arithmetic.  `resize_gray_reference` is the same arithmetic in NumPy, used by the CPU tests to pin the
tables against cv2 without a GPU."""
from __future__ import annotations

import functools
from typing import Tuple

import numpy as np


@functools.lru_cache(maxsize=64)
def _axis_table(n_dst: int, n_src: int, clamp: bool) -> Tuple[np.ndarray, np.ndarray]:
    """OpenCV resizeLinear coefficient table along one axis: (index int32[n], coeffs int16[n,2]).
    Columns (`clamp=True`): an index outside [0, n_src-1] is clamped and its fraction zeroed.  Rows: OpenCV keeps
    the fraction and clips the two source rows instead, so the index is left unclamped (the consumer clips)."""
    scale = 1.0 / (n_dst / n_src)                                  # double, like cv::resize
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)               # (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        low = s < 0
        f[low] = 0.0
        s[low] = 0
        high = s >= n_src - 1
        f[high] = 0.0
        s[high] = n_src - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)      # saturate_cast<short>(cvRound(.))
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    coef = np.clip(np.stack([a0, a1], axis=1), -32768, 32767).astype(np.int16)
    return s, coef


def resize_tables(h: int, w: int, nh: int, nw: int):
    xi, xa = _axis_table(nw, w, True)
    yi, ya = _axis_table(nh, h, False)
    return xi, xa, yi, ya


def resize_gray_reference(img: np.ndarray, nh: int, nw: int) -> Tuple[np.ndarray, np.ndarray]:
    """NumPy restatement of the kernel: (colour u8 (nh,nw,3), gray f32 (nh,nw)) == cv.resize / cvtColor / 255."""
    h, w = img.shape[:2]
    xi, xa, yi, ya = resize_tables(h, w, nh, nw)
    src = img.astype(np.int32)
    x1 = np.minimum(xi + 1, w - 1)
    y1 = np.clip(yi + 1, 0, h - 1)
    yi = np.clip(yi, 0, h - 1)
    a0, a1 = xa[:, 0].astype(np.int32)[None, :, None], xa[:, 1].astype(np.int32)[None, :, None]
    b0, b1 = ya[:, 0].astype(np.int32)[:, None, None], ya[:, 1].astype(np.int32)[:, None, None]
    r0 = src[yi][:, xi] * a0 + src[yi][:, x1] * a1
    r1 = src[y1][:, xi] * a0 + src[y1][:, x1] * a1
    v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    color = np.clip(v, 0, 255).astype(np.uint8)
    c = color.astype(np.int32)
    g = (c[..., 0] * 3735 + c[..., 1] * 19235 + c[..., 2] * 9798 + (1 << 14)) >> 15
    return color, g.astype(np.uint8).astype(np.float32) / 255.0
