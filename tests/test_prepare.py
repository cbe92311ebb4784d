"""CPU: the fixed-point tables / arithmetic of the device `_prepare_images` against OpenCV itself
(cv.resize INTER_LINEAR + cv.cvtColor BGR2GRAY + /255, reference mvs_patchmatch.py:176-177)."""
import cv2 as cv
import numpy as np
import pytest

from b200dense.prepare import resize_gray_reference


@pytest.mark.parametrize('h,w,scale', [(96, 128, 0.5), (97, 131, 0.25), (300, 400, 0.25), (480, 640, 0.37),
                                       (72, 96, 1.0), (600, 800, 0.5), (50, 70, 1.7), (33, 17, 0.9)])
def test_resize_and_gray_match_opencv_bitexact(h, w, scale):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    nh, nw = int(h * scale), int(w * scale)
    ref_color = cv.resize(img, (nw, nh))
    ref_gray = cv.cvtColor(ref_color, cv.COLOR_BGR2GRAY).astype(np.float32) / 255.0
    color, gray = resize_gray_reference(img, nh, nw)
    assert np.array_equal(color, ref_color)
    assert np.array_equal(gray, ref_gray)


def test_gray_coefficients_exhaustive():
    b, g, r = np.meshgrid(np.arange(256), np.arange(256), np.arange(0, 256, 5), indexing='ij')
    full = np.ascontiguousarray(np.stack([b, g, r], -1).astype(np.uint8).reshape(256 * 256, -1, 3))
    ref = cv.cvtColor(full, cv.COLOR_BGR2GRAY).astype(np.int64)
    mine = (full[..., 0].astype(np.int64) * 3735 + full[..., 1].astype(np.int64) * 19235
            + full[..., 2].astype(np.int64) * 9798 + (1 << 14)) >> 15
    assert np.array_equal(mine, ref)
