#!/usr/bin/env python
"""Benchmark of the dense hot path (BASELINE.json metric: PatchMatch views/s @1600x1200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (per GPU, weak scaling): PatchMatch MVS of V=8 reference views of a 64-view synthetic
turntable at 1600x1200, S=5 source views (the reference's own `_select_source_views`), 11x11 NCC,
3 iterations, 8 samples (E=31 cost evaluations per view), followed by the conf>=3 mask and
back-projection of every view -- the per-GPU share of BASELINE.json configs[3].  One "step" = those
V views.  Arithmetic: the bit-exact parity kernels (the ones the parity tests gate).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM), `e2e` goes
through the reference-facing `PatchMatchMVS.compute_depth_maps` (batch form of `_patchmatch_cuda`) with pinned
HOST buffers: H2D of every gray the step touches and D2H of depth/normal/confidence of every view inside the
timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, '3d-reconstruction-tool_b200')
for _p in (ROOT, PKG, os.path.join(PKG, 'src')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_SCENE_VIEWS = 64
METRIC = 'patchmatch_views_per_s_1600x1200'   # renamed below when --width/--height are overridden
UNIT = 'views/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--width', type=int, default=1600)
    ap.add_argument('--height', type=int, default=1200)
    ap.add_argument('--views-per-gpu', type=int, default=8)
    ap.add_argument('--patch', type=int, default=11)
    ap.add_argument('--src', type=int, default=5)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--samples', type=int, default=8)
    ap.add_argument('--streams', type=int, default=4, help='independent views in flight per GPU')
    ap.add_argument('--sweep-streams', type=int, default=2, help='plane-sweep reference views in flight per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-plane-sweep', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def workload_name(a):
    return (f"PatchMatch MVS {a.width}x{a.height}, {a.views_per_gpu} ref views/GPU of a {N_SCENE_VIEWS}-view synthetic "
            f"turntable, S={a.src}, {a.patch}x{a.patch} NCC, {a.iters} iters x (2 shifts + {a.samples} samples) + confidence, "
            f"then conf>=3 mask + back-projection (BASELINE configs[3] per-GPU share)")


def scene_for(a, ref_views, device):
    """Poses for all 64 views; pixels only for the views this rank touches."""
    from b200dense.scenes import make_scene
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    sc0 = make_scene('turntable', N_SCENE_VIEWS, a.width, a.height, arc_deg=360.0 / N_SCENE_VIEWS, seed=0,
                     device='cpu', render=[])
    poses = {i: CameraPose(R=sc0.poses[i][0], t=sc0.poses[i][1]) for i in sc0.poses}
    cam = Camera(K=sc0.K.copy(), dist=np.zeros(5))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        pm = PatchMatchMVS(cam, scale=1.0, patch_size=a.patch, num_iterations=a.iters, num_samples=a.samples,
                           min_views=3, num_src_views=a.src)
    pm.depth_min, pm.depth_max = sc0.depth_range
    srcs = {r: pm._select_source_views(r, sorted(poses), poses, k=a.src) for r in ref_views}
    need = sorted(set(ref_views) | {s for r in ref_views for s in srcs[r]})
    sc = make_scene('turntable', N_SCENE_VIEWS, a.width, a.height, arc_deg=360.0 / N_SCENE_VIEWS, seed=0,
                    device=device, render=need)
    return sc, poses, pm, srcs, need


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def oracle_view_runner(a, sc, poses, srcs, ref):
    """Closure running the CPU oracle on one full view of the workload (injected RNG)."""
    from b200dense.engine import kinv32
    from b200dense.scenes import gray_f32
    from oracle import oracle as orc
    gray = gray_f32(sc)
    H, W = a.height, a.width
    rng = np.random.default_rng(0)
    n = a.iters * a.samples
    dmin, dmax = sc.depth_range
    depth0 = np.exp(rng.random((H, W), dtype=np.float32) * np.float32(np.log(dmax) - np.log(dmin)) + np.float32(np.log(dmin)))
    n0 = rng.standard_normal((H, W, 3), dtype=np.float32) * np.float32(0.3)
    n0[..., 2] = -1
    normal0 = n0 / np.linalg.norm(n0, axis=-1, keepdims=True)
    ru = rng.random((n, H, W), dtype=np.float32)
    rn = rng.standard_normal((n, H, W, 3), dtype=np.float32)
    s = srcs[ref]
    cams = orc.Cams(sc.K, kinv32(sc.K), poses[ref].R, poses[ref].t, [poses[i].R for i in s], [poses[i].t for i in s])

    def run():
        return orc.pm_view(gray[ref], [gray[i] for i in s], cams, a.patch, a.iters, a.samples, dmin, dmax,
                           depth0, normal0, ru, rn)
    return run, orc.num_threads()


def run_reference(a):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port, all host threads)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ref = 0
    sc, poses, pm, srcs, _ = scene_for(a, [ref], 'cpu')
    run, threads = oracle_view_runner(a, sc, poses, srcs, ref)
    for _ in range(min(a.warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    dt = time.perf_counter() - t0
    v = a.steps / dt
    sample = f'1 view per step ({a.width}x{a.height}, S={a.src}, k={a.patch}, E={a.iters * (2 + a.samples) + 1}) instead of {a.views_per_gpu}'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': min(a.warmup, 1), 'ms_per_step': dt / a.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(a), 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    global METRIC
    a = parse()
    METRIC = f'patchmatch_views_per_s_{a.width}x{a.height}'
    if a.impl == 'reference':
        return run_reference(a)

    import torch
    import torch.distributed as dist
    from b200dense import dist as bdist
    from b200dense.engine import DenseEngine, kinv32, make_cams
    from b200dense.scenes import gray_f32

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the dense path)')
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry
        # exactly one JSON line, so route fd 1 to stderr while the communicator comes up.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device(dev))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    eng = DenseEngine(dev)

    V, H, W = a.views_per_gpu, a.height, a.width
    ref_views = list(range(N_SCENE_VIEWS))[rank::world][:V]
    sc, poses, pm, srcs, need = scene_for(a, ref_views, dev)
    gray = gray_f32(sc)
    d_gray = {i: eng.to_device(gray[i]) for i in need}
    d_bgr = {r: eng.to_device(sc.images[r]['image'], torch.uint8) for r in ref_views}
    Kinv32, Kinv64 = kinv32(sc.K), np.linalg.inv(sc.K)
    cams = {r: make_cams(sc.K, Kinv32, poses[r].R, poses[r].t, [poses[i].R for i in srcs[r]], [poses[i].t for i in srcs[r]])
            for r in ref_views}
    dmin, dmax = sc.depth_range
    outs = {r: dict(depth=torch.empty((H, W), device=dev), normal=torch.empty((H, W, 3), device=dev),
                    confidence=torch.empty((H, W), device=dev)) for r in ref_views}
    NS = max(1, min(a.streams, V))
    pts_buf = {r: torch.empty((H * W, 3), dtype=torch.float64, device=dev) for r in ref_views}
    col_buf = {r: torch.empty((H * W, 3), dtype=torch.uint8, device=dev) for r in ref_views}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    all_views = sorted(v for rk in range(world) for v in list(range(N_SCENE_VIEWS))[rk::world][:V])

    def view_job(r):
        def job(slot):
            o = eng.pm_view(d_gray[r], [d_gray[i] for i in srcs[r]], cams[r], a.patch, a.iters, a.samples, dmin, dmax,
                            seed=1234, view_id=r, out=outs[r], slot=slot)
            return eng.mask_backproject(o['depth'], o['confidence'], d_bgr[r], Kinv64, poses[r].R, poses[r].t, 3.0, 0,
                                        points=pts_buf[r], colors=col_buf[r], sync=False, slot=slot)
        return job
    jobs = [view_job(r) for r in ref_views]
    gathered = {}

    def step():
        res = eng.run_views(jobs, NS)
        if world > 1:
            # the job's only exchange: fused clouds of all ranks' views -> rank 0, sorted-view order (NCCL p2p)
            local = {}
            for r, (p, c, cnt) in zip(ref_views, res):
                m = int(cnt.item())
                local[r] = (p[:m], c[:m])
            gp, gc = bdist.gather_point_clouds(all_views, local, torch.device(dev))
            gathered['points'] = int(gp.shape[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    for e0, e1 in ev:
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        e0.record()
        step()
        e1.record()
    barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = V * world * a.steps / (ms / 1e3)

    # ---- per-kernel durations: one more step right after the timed region, views serialised on one stream
    # (with several views in flight the kernels of different views overlap and a per-launch time is not
    # attributable), CUDA events on the launching stream around every launch (b200d_profile_begin/_end).
    flush.zero_()
    eng.profile_begin(65536)
    eng.run_views(jobs, 1)
    torch.cuda.synchronize()
    kms, ktag = eng.profile_end()
    names = {400 + a.patch: 'ncc_kernel<REFINE>', 300 + a.patch: 'ncc_kernel<SHIFT>', 200 + a.patch: 'ncc_kernel<CONF>',
             a.patch: 'ncc_kernel<STATS>', 2020: 'sample_kernel<REFINE>', 2010: 'sample_kernel<SHIFT>',
             2001: 'sample_kernel<MAP,conf>', 3000: 'pm_init_kernel', 4000: 'fuse kernels (3)'}
    tot_k = float(kms.sum())
    kernels = {names.get(int(tg), str(int(tg))): {'launches_per_step': int((ktag == tg).sum()),
                                                   'avg_ms': float(kms[ktag == tg].mean()),
                                                   'share_of_step': float(kms[ktag == tg].sum() / tot_k)}
               for tg in np.unique(ktag)}
    dom_ms = kernels['ncc_kernel<REFINE>']['avg_ms']
    smp_ms = kernels['sample_kernel<REFINE>']['avg_ms']
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks else '6650 GB/s (of fallback)'
    # algorithmic HBM bytes of one cost evaluation = images 4(1+S) + state 2*(4+12+4) per pixel.
    # One evaluation = sample_kernel + ncc_kernel; the sampled images in between are L2-resident scratch.
    alg_bytes = (44 + 4 * a.src) * H * W
    ach = alg_bytes / (dom_ms * 1e-3) / 1e9
    evals = H * W * a.src                                     # (pixel, source) evaluations per launch
    fma_peak_nominal = 148 * 128 * peaks.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
    FMA_PEAK_MEASURED = 37.1                                  # T lane-FMA/s: best measured FFMA2 rate on this pool
                                                              # (profiles/microbench/ffma2_issue.cu: 37.08; ffma2.cu: 36.3)
    pair_s = (dom_ms + smp_ms) * 1e-3
    fp32 = {
        'note': 'the evaluation is FP32-pipe bound, not HBM bound (SURVEY 8d); lane-ops count an FMA as one; '
                'parity arithmetic needs 3k^2 FMAs per (pixel, source) by construction',
        'eval_sources_per_s_ncc_kernel': evals / (dom_ms * 1e-3),
        'eval_sources_per_s_sample_plus_ncc': evals / pair_s,
        'peak_tlaneops_nominal': fma_peak_nominal, 'peak_tlaneops_measured_ffma': FMA_PEAK_MEASURED,
        'ncc_kernel_fma_tlaneops': evals * (3 * a.patch * a.patch + 2 * (a.patch + 3) * 14 / 8 + 20) / (dom_ms * 1e-3) / 1e12,
        'pair_tlaneops_parity_model': evals * (40 + 3 * a.patch * a.patch) / pair_s / 1e12,
        'pair_tlaneops_fast_model': evals * (40 + 6 * (a.patch - 1)) / pair_s / 1e12,
    }
    fp32['ncc_kernel_frac_of_measured_fma_peak'] = fp32['ncc_kernel_fma_tlaneops'] / FMA_PEAK_MEASURED
    fp32['pair_frac_parity_model'] = fp32['pair_tlaneops_parity_model'] / FMA_PEAK_MEASURED
    fp32['pair_frac_fast_model'] = fp32['pair_tlaneops_fast_model'] / FMA_PEAK_MEASURED
    # Headline roofline of the dominant kernel against the resource that binds it: the FP32 FMA pipe.
    # Algorithmic flops of ncc_kernel per (pixel, source): three k x k window sums as FMAs (2 flops each), the two
    # products s*s and r*s, ~20 for variance / covariance / NCC / aggregation.  Peak = measured FFMA/FFMA2 rate
    # (profiles/microbench/ffma2_issue.cu: 37.1 T lane-FMA/s = 74.2 TFLOP/s at 1965 MHz; nominal 148*128*2*1.965 = 74.5).
    ncc_flops = evals * (6 * a.patch * a.patch + 22)
    ach_tf = ncc_flops / (dom_ms * 1e-3) / 1e12
    roofline = {'bound': 'fp32', 'achieved': ach_tf, 'peak': 2 * FMA_PEAK_MEASURED, 'unit': 'TFLOP/s',
                'frac': ach_tf / (2 * FMA_PEAK_MEASURED),
                'traffic': 97.2e6, 'traffic_source': 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch '
                                                     '(profiles/r01_ncu_full_ncc_sample.txt)',
                'kernel': f'ncc_kernel<{a.patch},REFINE>', 'launch_ms': dom_ms,
                'launches_timed': kernels['ncc_kernel<REFINE>']['launches_per_step'],
                'kernel_share_of_step': kernels['ncc_kernel<REFINE>']['share_of_step'],
                'algorithmic_flops_per_launch': ncc_flops,
                'peak_source': 'measured FFMA2 rate on this pool (profiles/microbench/ffma2_issue.cu), of measured',
                'how': 'CUDA events on the launching stream around every launch, one serialised step right after the '
                       'timed region',
                'note': 'contract lists hbm|tensor; this path is neither: no dense contraction (no tensor cores) and '
                        '10 % of HBM peak -- the FP32 pipe binds it, so the fraction is quoted against the measured FFMA peak',
                'hbm': {'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                        'algorithmic_bytes_per_launch': alg_bytes, 'peak_source': peak_src},
                'kernels': kernels, 'fp32': fp32}

    # ---- end to end through the reference-facing API with pinned host buffers
    e2e = None
    if not a.no_e2e:
        pm.cache_images = False
        pm.seed = 1234
        pinned = {i: torch.from_numpy(gray[i]).pin_memory() for i in need}
        proc = {i: {'gray': pinned[i].numpy(), 'shape': (H, W)} for i in need}

        def step_e2e():
            maps = pm.compute_depth_maps(proc, poses, ref_views, srcs)
            assert len(maps) == V and maps[ref_views[0]].depth.shape == (H, W)
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ts = time.perf_counter()
            step_e2e()
            print(f'[e2e] step {1e3 * (time.perf_counter() - ts):.1f} ms', file=sys.stderr)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {'value': V * world * a.steps / float(dt.item()), 'unit': UNIT,
               'h2d_bytes_per_step': len(need) * H * W * 4, 'd2h_bytes_per_step': V * 5 * H * W * 4,
               'api': 'core.mvs_patchmatch.PatchMatchMVS.compute_depth_maps (batch form of _patchmatch_cuda: pinned host '
                      'grays in, every distinct image uploaded once per step, host depth/normal/confidence out)'}

    # ---- secondary metric: plane sweep Gpix*planes/s (BASELINE configs[2] shape, one view per GPU)
    sweep = None
    if not a.no_plane_sweep:
        from b200dense.scenes import make_scene
        Ws, Hs, D, N, k = 1920, 1080, 256, 6, 7
        nb_views = list(range(0, 10))
        scs = make_scene('turntable', N_SCENE_VIEWS, Ws, Hs, arc_deg=360.0 / N_SCENE_VIEWS, seed=0, device=dev, render=nb_views)
        gs = gray_f32(scs)
        dg = {i: eng.to_device(gs[i]) for i in nb_views}
        refs_s = [3, 4, 5, 6]                                  # 4 reference views per rep, 2 in flight
        depths = (1.0 / np.linspace(1 / scs.depth_range[1], 1 / scs.depth_range[0], D)).astype(np.float32)
        Ki = kinv32(scs.K)

        def sweep_job(r):
            nbs = [r - 1, r + 1, r - 2, r + 2, r - 3, r + 3]   # 6 nearest cameras on the circle
            cs = make_cams(scs.K, Ki, scs.poses[r][0], scs.poses[r][1], [scs.poses[i][0] for i in nbs],
                           [scs.poses[i][1] for i in nbs])
            return lambda slot: eng.plane_sweep(dg[r], [dg[i] for i in nbs], cs, depths, k, 0.8, slot=slot)
        sjobs = [sweep_job(r) for r in refs_s]
        eng.run_views(sjobs[:2], a.sweep_streams)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_views(sjobs, a.sweep_streams)
        e1.record()
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        gpp = world * len(refs_s) * Hs * Ws * D / (float(tms.item()) / 1e3) / 1e9
        sweep = {'metric': 'plane_sweep_gpix_planes_per_s', 'value': gpp, 'unit': 'Gpix*planes/s',
                 'config': f'{Ws}x{Hs}, D={D}, N={N}, {k}x{k} NCC, {len(refs_s)} ref views per GPU ({a.sweep_streams} in flight), '
                           f'parity arithmetic (BASELINE configs[2] shape)',
                 'ms_per_view': float(tms.item()) / len(refs_s),
                 'evals_per_s': gpp * 1e9 * N,
                 'fp32_frac_parity_model': gpp * 1e9 * N * (40 + 3 * k * k) / world / (FMA_PEAK_MEASURED * 1e12),
                 'fp32_frac_fast_model': gpp * 1e9 * N * (40 + 6 * (k - 1)) / world / (FMA_PEAK_MEASURED * 1e12)}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only): the oracle on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        run, threads = oracle_view_runner(a, sc, poses, srcs, ref_views[0])
        t0 = time.perf_counter()
        n = 0
        while n < 3 and (time.perf_counter() - t0) < 12.0:
            run()
            n += 1
        dt = time.perf_counter() - t0
        cpu = {'value': n / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'{n} full view(s) of the same workload ({W}x{H}, S={a.src}, k={a.patch}, E={a.iters * (2 + a.samples) + 1}) '
                         f'through oracle/dense_oracle.c (OpenMP, {threads} threads), {dt:.1f} s'}

    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(a), 'views_per_step_per_gpu': V, 'arithmetic': 'parity (bit-exact vs oracle)',
                       'rng': 'philox4x32-10 in-kernel', 'l2': 'flushed between timed steps (512 MiB write)',
                       'views_in_flight': NS,
                       'parallelism': f'view-sharded dp{world}' + (', fused clouds gathered to rank 0 every step (NCCL p2p)'
                                                                   if world > 1 else '')},
            'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
            'plane_sweep': sweep}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
