"""Pinhole camera and pose carriers with the reference's field names and conventions
(reference: src/core/camera.py:10-139): K 3x3 intrinsics, dist = [k1,k2,p1,p2,k3];
pose maps world to camera, X_cam = R @ X_world + t; calibration npz keys `mtx`, `dist`."""
from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class Camera:
    K: np.ndarray
    dist: np.ndarray

    fx = property(lambda self: self.K[0, 0])
    fy = property(lambda self: self.K[1, 1])
    cx = property(lambda self: self.K[0, 2])
    cy = property(lambda self: self.K[1, 2])

    def project(self, points_3d: np.ndarray) -> np.ndarray:
        """Camera-frame points (N,3) -> pixels (N,2)."""
        xy = points_3d[:, :2] / points_3d[:, 2:3]
        return np.column_stack([self.fx * xy[:, 0] + self.cx, self.fy * xy[:, 1] + self.cy])

    def unproject(self, points_2d: np.ndarray, depth: float = 1.0) -> np.ndarray:
        """Pixels (N,2) -> camera-frame points at the given depth (N,3)."""
        xn = (points_2d[:, 0] - self.cx) / self.fx
        yn = (points_2d[:, 1] - self.cy) / self.fy
        return np.column_stack([xn * depth, yn * depth, np.full(len(points_2d), 1.0) * depth])


@dataclass
class CameraPose:
    R: np.ndarray
    t: np.ndarray

    @property
    def center(self) -> np.ndarray:
        return -self.R.T @ self.t.ravel()

    @property
    def projection_matrix(self) -> np.ndarray:
        return np.hstack([self.R, self.t.reshape(3, 1)])

    def transform_points(self, points_world: np.ndarray) -> np.ndarray:
        return (self.R @ points_world.T).T + self.t.ravel()

    @staticmethod
    def identity() -> 'CameraPose':
        return CameraPose(R=np.eye(3), t=np.zeros(3))


def load_calibration(calibration_path: str) -> Camera:
    """Camera from a calibration .npz (keys `mtx`, `dist`; fewer than five distortion coefficients are zero-padded).
    Same banner lines and the same FileNotFoundError text as the reference (src/core/camera.py:111-138)."""
    src = Path(calibration_path)
    try:
        with np.load(str(src)) as npz:
            mtx = np.asarray(npz['mtx'], dtype=np.float64)
            coeffs = np.asarray(npz['dist'], dtype=np.float64).reshape(-1)
    except FileNotFoundError:
        raise FileNotFoundError(f"Calibration file not found: {src}") from None
    cam = Camera(K=mtx, dist=np.concatenate([coeffs, np.zeros(max(0, 5 - coeffs.size))]))
    for line in (f"Loaded calibration from {src.name}",
                 f"  Focal length: fx={cam.fx:.1f}, fy={cam.fy:.1f}",
                 f"  Principal point: cx={cam.cx:.1f}, cy={cam.cy:.1f}"):
        print(line)
    return cam
