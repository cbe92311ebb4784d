"""Point-cloud output in the reference's format (reference: src/core/utils.py:8-87):
ASCII PLY, `x y z` as %.6f floats and `red green blue` as uchar, one vertex per line.
The vertex lines are written by a buffered C routine of the dense library (b200d_write_ply_ascii, same bytes as
the reference's per-point Python loop); if the library has not been built a pure-Python writer is used."""
from pathlib import Path

import numpy as np

_HEADER = ("ply\nformat ascii 1.0\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\n"
           "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")


def _native_writer():
    try:
        from b200dense import _cabi
        return _cabi.load().b200d_write_ply_ascii
    except Exception:                       # library not built / not importable: formatting is host-only work
        return None


def save_ply(points: np.ndarray, colors: np.ndarray, output_path: str):
    out = Path(output_path)
    out.parent.mkdir(parents=True, exist_ok=True)
    n = len(points)
    native = _native_writer()
    if native is not None:
        import ctypes as C
        xyz = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(n, 3))
        rgb = np.ascontiguousarray(np.asarray(colors).reshape(n, 3).astype(int).astype(np.uint8))
        rc = native(str(out).encode(), xyz.ctypes.data_as(C.POINTER(C.c_double)),
                    rgb.ctypes.data_as(C.POINTER(C.c_uint8)), n)
        if rc != 0:
            raise OSError(f"could not write {out}")
        print(f"Saved {n:,} points to {out}")
        return
    with open(out, 'w') as f:
        f.write(_HEADER.format(n=n))
        if n:
            xyz = np.asarray(points, dtype=np.float64).reshape(n, 3)
            rgb = np.asarray(colors).reshape(n, 3).astype(int)
            block = 1 << 18
            for s in range(0, n, block):
                p, c = xyz[s:s + block], rgb[s:s + block]
                rows = ["%.6f %.6f %.6f %d %d %d" % (a, b, d, r, g, h)
                        for (a, b, d), (r, g, h) in zip(p.tolist(), c.tolist())]
                f.write("\n".join(rows))
                f.write("\n")
    print(f"Saved {n:,} points to {out}")


def save_cameras_ply(poses: dict, output_path: str, scale: float = 0.5):
    """Two vertices per camera: centre (red) and centre + scale * (-R[2]) (green)."""
    pts, cols = [], []
    for pose in poses.values():
        c = pose.center
        pts += [c, c - pose.R[2, :] * scale]
        cols += [[255, 0, 0], [0, 255, 0]]
    save_ply(np.array(pts), np.array(cols, dtype=np.uint8), str(Path(output_path)))


def compute_scene_bounds(points: np.ndarray) -> dict:
    if len(points) == 0:
        z = np.zeros(3)
        return {'min': z, 'max': z.copy(), 'center': z.copy(), 'size': 0}
    lo, hi = points.min(axis=0), points.max(axis=0)
    return {'min': lo, 'max': hi, 'center': (lo + hi) / 2, 'size': np.linalg.norm(hi - lo)}
