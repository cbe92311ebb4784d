"""Shared helpers for the parity tests (fixture loading, RNG replay, metrics)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_cams(orc, g, ref, srcs):
    return orc.Cams(g['K_scaled'], g['K_inv32'], g['R'][ref], g['t'][ref], [g['R'][i] for i in srcs],
                    [g['t'][i] for i in srcs])


def engine_cams(g, ref, srcs):
    from b200dense.engine import make_cams
    return make_cams(g['K_scaled'], g['K_inv32'], g['R'][ref], g['t'][ref], [g['R'][i] for i in srcs],
                     [g['t'][i] for i in srcs])


def replay_rng(g):
    """Re-draw the reference's RNG stream (torch CPU generator, call order of
    mvs_patchmatch.py:271,279-280,471,475) and the initial hypotheses (:268-281).
    Returns None if this host's torch does not reproduce the recorded stream."""
    import torch
    import torch.nn.functional as F
    H, W = g['gray'][0].shape
    iters, samples = int(g['iters']), int(g['samples'])
    torch.manual_seed(int(g['seed']))
    iu, nx, ny = torch.rand(H, W), torch.randn(H, W), torch.randn(H, W)
    ru, rn = [], []
    for _ in range(iters * samples):
        ru.append(torch.rand(H, W))
        rn.append(torch.randn(H, W, 3))
    ru = torch.stack(ru).numpy()
    rn = torch.stack(rn).numpy()
    got = [sha(iu.numpy()), sha(nx.numpy()), sha(ny.numpy()), sha(ru), sha(rn)]
    if got != list(g['rng_sha']):
        return None
    dmin, dmax = float(g['depth_min']), float(g['depth_max'])
    ld0, ld1 = np.log(dmin), np.log(dmax)
    depth0 = torch.exp(iu * (ld1 - ld0) + ld0).numpy()
    n0 = torch.zeros(H, W, 3)
    n0[:, :, 2] = -1
    n0[:, :, 0] = nx * 0.3
    n0[:, :, 1] = ny * 0.3
    normal0 = F.normalize(n0, dim=-1).numpy()
    return dict(depth0=depth0, normal0=normal0, ref_u=ru, ref_n=rn)


def normal_angle_deg(a, b):
    """Angle between unit-normal maps; identical vectors (incl. the zero border normals) count as 0."""
    dot = np.clip((a * b).sum(-1), -1.0, 1.0)
    ang = np.degrees(np.arccos(dot))
    same = np.all(a == b, axis=-1)
    return np.where(same, 0.0, ang)


def pm_agreement(depth, normal, g_depth, g_normal):
    """north_star metric: fraction of pixels within 0.5 % relative depth AND 1 degree normal error."""
    rel = np.abs(depth - g_depth) / np.abs(g_depth)
    ang = normal_angle_deg(normal, g_normal)
    return float(((rel <= 0.005) & (ang <= 1.0)).mean())
