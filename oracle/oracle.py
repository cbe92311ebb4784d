"""ctypes front-end of the CPU oracle (oracle/dense_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'liborc.so')
_lib = None

_f = np.float32


def build(force: bool = False) -> str:
    """Compile liborc.so with the recipe in oracle/Makefile (gcc, no CUDA)."""
    src = os.path.join(_HERE, 'dense_oracle.c')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_mask_backproject.restype = C.c_int64
    return _lib


def _p(a, ty=C.c_float):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=_f)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def box_mean(x, k):
    x = _f32(x)
    out = np.empty_like(x)
    lib().orc_box_mean(_p(x), x.shape[0], x.shape[1], int(k), _p(out))
    return out


class Cams:
    """fp32 camera block exactly as the reference uploads it (mvs_patchmatch.py:237-257,
    dense_stereo.py:232-259): K_scaled, R, t cast f64->f32; Kinv = fp32 inverse supplied by caller."""

    def __init__(self, K_scaled, Kinv32, R_ref, t_ref, R_src, t_src):
        self.K = _f32(K_scaled)
        self.Kinv = _f32(Kinv32)
        self.Rr = _f32(R_ref)
        self.tr = _f32(t_ref)
        self.Rs = _f32(np.stack(R_src))
        self.ts = _f32(np.stack(t_src))
        self.S = self.Rs.shape[0]

    def args(self):
        return _p(self.K), _p(self.Kinv), _p(self.Rr), _p(self.tr), _p(self.Rs), _p(self.ts)


def warp_sample(src, cams: Cams, depth):
    """(sampled, u, v, z) of source 0 of ``cams`` at a per-pixel depth map."""
    src = _f32(src)
    depth = _f32(depth)
    H, W = src.shape
    o = [np.empty((H, W), _f) for _ in range(4)]
    lib().orc_warp_sample(_p(src), H, W, *cams.args(), _p(depth), *[_p(a) for a in o])
    return tuple(o)


def plane_sweep(ref, nbrs, cams: Cams, depths, k, thresh, dump_ncc=False):
    ref = _f32(ref)
    nb = _f32(np.stack(nbrs))
    H, W = ref.shape
    N = nb.shape[0]
    d32 = _f32(depths)
    D = d32.shape[0]
    depth_map = np.empty((H, W), _f)
    conf = np.empty((H, W), _f)
    idx = np.empty((H, W), np.int32)
    ncc = np.empty((D, N, H, W), _f) if dump_ncc else None
    lib().orc_plane_sweep(_p(ref), _p(nb), N, H, W, *cams.args(), _p(d32), D, int(k), C.c_float(thresh),
                          _p(depth_map), _p(conf), _p(idx, C.c_int32), _p(ncc))
    return dict(depth_map=depth_map, confidence=conf, best_idx=idx, ncc=ncc)


def pm_cost(ref, srcs, cams: Cams, depth, k, mode=0):
    ref = _f32(ref)
    sr = _f32(np.stack(srcs))
    H, W = ref.shape
    depth = _f32(depth)
    out = np.empty((H, W), _f)
    lib().orc_pm_cost(_p(ref), _p(sr), sr.shape[0], H, W, *cams.args(), _p(depth), int(k), int(mode), _p(out))
    return out


def range_schedule(depth_min, depth_max, iters):
    """fp32 casts of the Python-double expressions at mvs_patchmatch.py:466-467."""
    dr = np.array([(depth_max - depth_min) * (0.5 ** it) for it in range(iters)], dtype=np.float64).astype(_f)
    nr = np.array([0.5 * (0.5 ** it) for it in range(iters)], dtype=np.float64).astype(_f)
    return dr, nr


def pm_view(ref, srcs, cams: Cams, k, iters, samples, depth_min, depth_max, depth0, normal0, rand_u, rand_n,
            dump_costs=False):
    ref = _f32(ref)
    sr = _f32(np.stack(srcs))
    H, W = ref.shape
    depth = _f32(depth0).copy()
    normal = _f32(normal0).copy()
    ru = _f32(rand_u)
    rn = _f32(rand_n)
    assert ru.shape == (iters * samples, H, W) and rn.shape == (iters * samples, H, W, 3)
    dr, nr = range_schedule(depth_min, depth_max, iters)
    conf = np.empty((H, W), _f)
    best = np.empty((H, W), _f)
    costs = np.empty((iters * (2 + samples), H, W), _f) if dump_costs else None
    lib().orc_pm_view(_p(ref), _p(sr), sr.shape[0], H, W, *cams.args(), int(k), int(iters), int(samples),
                      C.c_float(_f(depth_min)), C.c_float(_f(depth_max)), _p(dr), _p(nr), _p(ru), _p(rn),
                      _p(depth), _p(normal), _p(conf), _p(best), _p(costs))
    return dict(depth=depth, normal=normal, confidence=conf, best_cost=best, costs=costs)


def mask_backproject(depth, conf, bgr, Kinv64, R, t, min_conf, mode):
    depth = _f32(depth)
    conf = _f32(conf)
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    H, W = depth.shape
    Ki = np.ascontiguousarray(Kinv64, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64).ravel()
    d = C.POINTER(C.c_double)
    u8 = C.POINTER(C.c_uint8)
    m = lib().orc_mask_backproject(_p(depth), _p(conf), bgr.ctypes.data_as(u8), H, W,
                                   Ki.ctypes.data_as(d), R.ctypes.data_as(d), t.ctypes.data_as(d),
                                   C.c_float(min_conf), int(mode), None, None)
    pts = np.empty((m, 3), np.float64)
    col = np.empty((m, 3), np.uint8)
    if m:
        lib().orc_mask_backproject(_p(depth), _p(conf), bgr.ctypes.data_as(u8), H, W,
                                   Ki.ctypes.data_as(d), R.ctypes.data_as(d), t.ctypes.data_as(d),
                                   C.c_float(min_conf), int(mode), pts.ctypes.data_as(d), col.ctypes.data_as(u8))
    return pts, col


def geo_consistency(depth_ref, depth_srcs, cams: Cams, px_thresh, rel_thresh):
    """EXTENSION definition (no dackey-wav counterpart): number of sources whose depth map agrees with depth_ref."""
    depth_ref = _f32(depth_ref)
    ds = _f32(np.stack(depth_srcs))
    H, W = depth_ref.shape
    out = np.empty((H, W), _f)
    lib().orc_geo_consistency(_p(depth_ref), _p(ds), ds.shape[0], H, W, *cams.args(), C.c_float(px_thresh),
                              C.c_float(rel_thresh), _p(out))
    return out


def viewprop_candidates(depth, normal, depth_src, normal_src, s, cams: Cams, depth_min, depth_max):
    """EXTENSION definition: view-propagation proposal (cand_depth, cand_normal, allow) from source `s` of `cams`."""
    depth, normal = _f32(depth), _f32(normal)
    depth_src, normal_src = _f32(depth_src), _f32(normal_src)
    H, W = depth.shape
    cd = np.empty((H, W), _f)
    cn = np.empty((H, W, 3), _f)
    al = np.empty((H, W), np.uint8)
    lib().orc_viewprop_candidates(_p(depth), _p(normal), _p(depth_src), _p(normal_src), int(s), H, W, *cams.args(),
                                  C.c_float(_f(depth_min)), C.c_float(_f(depth_max)), _p(cd), _p(cn), _p(al, C.c_uint8))
    return cd, cn, al.astype(bool)
