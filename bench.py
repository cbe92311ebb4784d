#!/usr/bin/env python
"""Benchmark of the dense hot path (BASELINE.json metric: PatchMatch views/s @1600x1200 and plane-sweep Gpix*planes/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--arith parity|fast]
                    [--consistency photometric|geometric] [--config default|full4k]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (per GPU, weak scaling): PatchMatch MVS of V=8 reference views of a 64-view synthetic turntable at 1600x1200,
S=5 source views (the reference's own `_select_source_views`), 11x11 NCC, 3 iterations, 8 samples (E=31 cost
evaluations per view), followed by the conf>=3 mask and back-projection of every view -- the per-GPU share of
BASELINE.json configs[3]; at N>1 the fused clouds of all ranks are gathered to rank 0 inside every step.  One "step" =
those V views.  Arithmetic: the bit-exact parity kernels (the ones the parity tests gate) unless --arith fast.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM); `e2e` goes through the
reference-facing `PatchMatchMVS.compute_depth_maps` (batch form of `_patchmatch_cuda`) with pinned HOST buffers: H2D of
every gray the step touches and D2H of depth / normal / confidence of every view inside the timed region.  Further legs
in the same line: `e2e_reconstruct` (PatchMatchMVS.reconstruct(): host photographs -> filtered host cloud),
`plane_sweep` (its own value / e2e / roofline / cpu_baseline), `exchange` (NCCL all-gather of depth + normal maps and
the geometric filter -- BASELINE configs[3]'s extension), `gather_check` (N>1: the cloud rank 0 holds is bit-checked
against a recomputation), `fast_mode` (the opt-in, not bit-exact arithmetic, with its measured disagreement).

`--config full4k`: BASELINE configs[4] end to end -- N-view 3840x2160 turntable through reconstruct() -> post-filter ->
save_ply, with the share of each phase (its own JSON line).
"""
import argparse
import contextlib
import io
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
# what the reference itself does on CPU (unmodified PyTorch path, SURVEY.md section 6, survey container, 8 vCPU): context
# for the cpu_baseline figures below, which time the C/OpenMP PORT of that path, not PyTorch
REFERENCE_PYTORCH_CPU = {'patchmatch_views_per_s_1280x960_S5_k11': 0.042, 'plane_sweep_gpix_planes_per_s_cfg0': 0.0053,
                         'hardware': '8 vCPU, torch 2.11 CPU (SURVEY.md section 6)',
                         'note': 'cpu_baseline / --impl reference time the C/OpenMP port of the same arithmetic (oracle/), '
                                 'which is ~17-25x faster than the reference\'s own PyTorch CPU path'}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--arith', default='parity', choices=['parity', 'fast'])
    ap.add_argument('--consistency', default='photometric', choices=['photometric', 'geometric'],
                    help='geometric: all-gather of depth+normal maps and the geometric filter inside every timed step')
    ap.add_argument('--config', default='default', choices=['default', 'full4k'])
    ap.add_argument('--width', type=int, default=1600)
    ap.add_argument('--height', type=int, default=1200)
    ap.add_argument('--views-per-gpu', type=int, default=8)
    ap.add_argument('--patch', type=int, default=11)
    ap.add_argument('--src', type=int, default=5)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--samples', type=int, default=8)
    ap.add_argument('--streams', type=int, default=4, help='independent views in flight per GPU')
    ap.add_argument('--sweep-streams', type=int, default=2, help='plane-sweep reference views in flight per GPU')
    ap.add_argument('--full4k-views', type=int, default=200)
    ap.add_argument('--ply-out', default=None, help='full4k: where to write the PLY (default: a temporary file)')
    ap.add_argument('--full4k-min-views', type=int, default=3,
                    help='full4k: fusion threshold (reference default 3; 1 stresses fusion / filter / PLY with a large cloud)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-plane-sweep', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip e2e_reconstruct / exchange / fast_mode legs')
    return ap.parse_args()


def workload_name(a):
    return (f"PatchMatch MVS {a.width}x{a.height}, {a.views_per_gpu} ref views/GPU of a {N_SCENE_VIEWS}-view synthetic "
            f"turntable, S={a.src}, {a.patch}x{a.patch} NCC, {a.iters} iters x (2 shifts + {a.samples} samples) + confidence, "
            f"then conf>=3 mask + back-projection (BASELINE configs[3] per-GPU share)")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def config_dict(a, world):
    """The `config` object of the JSON line -- the same for the b200 arm and the reference (CPU) arm, which times a
    bounded sample of this workload (its sample is described in `cpu_baseline.sample`, not here)."""
    ns = max(1, min(a.streams, a.views_per_gpu))
    return {'workload': workload_name(a), 'views_per_step_per_gpu': a.views_per_gpu,
            'arithmetic': 'parity (bit-exact vs oracle)' if a.arith == 'parity'
                          else 'B200D_FAST (opt-in; NOT bit-exact, fails the parity gates: see fast_mode)',
            'rng': 'philox4x32-10 in-kernel (production generator: checked statistically and for determinism; '
                   'the bit-exact parity tests inject the reference\'s own draws, incl. at this 1600x1200 shape)',
            'l2': 'flushed between timed steps (512 MiB write)', 'views_in_flight': ns,
            'consistency': a.consistency,
            'parallelism': f'view-sharded dp{world}' + (' (contiguous blocks of views), fused clouds gathered to rank 0 '
                                                        'every step (peer stores over NVLink into the final buffer; '
                                                        'NCCL p2p where peer memory is unavailable: see gather_check)'
                                                        if world > 1 else '')}


def scene_for(a, ref_views, device, extra_views=()):
    """Poses for all 64 views; pixels only for the views this rank touches (+ `extra_views` and their sources)."""
    from b200dense.scenes import make_scene
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    sc0 = make_scene('turntable', N_SCENE_VIEWS, a.width, a.height, arc_deg=360.0 / N_SCENE_VIEWS, seed=0,
                     device='cpu', render=[])
    poses = {i: CameraPose(R=sc0.poses[i][0], t=sc0.poses[i][1]) for i in sc0.poses}
    cam = Camera(K=sc0.K.copy(), dist=np.zeros(5))
    with quiet():
        pm = PatchMatchMVS(cam, scale=1.0, patch_size=a.patch, num_iterations=a.iters, num_samples=a.samples,
                           min_views=3, num_src_views=a.src, arith=a.arith)
    pm.depth_min, pm.depth_max = sc0.depth_range
    views = list(ref_views) + [v for v in extra_views if v not in ref_views]
    srcs = {r: pm._select_source_views(r, sorted(poses), poses, k=a.src) for r in views}
    need = sorted(set(views) | {s for r in views for s in srcs[r]})
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


def oracle_threads():
    """All host threads for the CPU arm, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)."""
    from oracle import oracle as orc
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    orc.set_threads(n)
    return orc.num_threads()


def oracle_view_runner(a, sc, poses, srcs, ref):
    """Closure running the CPU oracle on one full view of the workload (injected RNG)."""
    from b200dense.engine import kinv32
    from b200dense.scenes import gray_f32
    from oracle import oracle as orc
    threads = oracle_threads()
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
    return run, threads


def run_reference(a):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port, all host threads)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ref = 0
    sc, poses, pm, srcs, _ = scene_for(a, [ref], 'cpu')
    run, threads = oracle_view_runner(a, sc, poses, srcs, ref)
    for _ in range(a.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    dt = time.perf_counter() - t0
    v = a.steps / dt
    sample = (f'1 view per step ({a.width}x{a.height}, S={a.src}, k={a.patch}, E={a.iters * (2 + a.samples) + 1}) instead of '
              f'{a.views_per_gpu}; rates are per view')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': dt / a.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_dict(a, max(1, a.gpus)),
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'reference_pytorch_cpu': REFERENCE_PYTORCH_CPU,
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def guarded(world, name, fn):
    """Secondary legs must not cost the headline line: on a single GPU a failing leg is reported as {'error': ...}.  With
    several ranks an exception has to propagate (a rank that skipped a leg would leave the others waiting in its
    collectives; the launcher then ends the whole job)."""
    if world > 1:
        return fn()
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {'error': f'{name} leg failed: {type(e).__name__}: {e}'}


def load_json(path):
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


def init_dist(world, dev):
    import torch
    import torch.distributed as dist
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


def main():
    global METRIC
    a = parse()
    METRIC = f'patchmatch_views_per_s_{a.width}x{a.height}'
    if a.impl == 'reference':
        return run_reference(a)
    if a.config == 'full4k':
        return run_full4k(a)

    import torch
    import torch.distributed as dist
    from b200dense import _cabi, dist as bdist
    from b200dense.engine import DenseEngine, kinv32, make_cams
    from b200dense.scenes import gray_f32

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the dense path)')
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    init_dist(world, dev)
    eng = DenseEngine(dev)
    ARITH = _cabi.FAST if a.arith == 'fast' else _cabi.PARITY

    V, H, W = a.views_per_gpu, a.height, a.width
    # weak scaling: V consecutive views of the turntable per GPU (the library's block sharding, b200dense/dist.py)
    assert V * world <= N_SCENE_VIEWS, 'views-per-gpu x gpus exceeds the synthetic turntable'
    all_views = list(range(V * world))
    ref_views = bdist.shard_views(all_views, rank, world)
    pos_of = {v: p for p, v in enumerate(all_views)}
    # gather check (N>1): rank 0 recomputes the first view owned by rank 1 -> it needs that view's images too
    check_view = bdist.shard_views(all_views, 1, world)[0] if world > 1 else None
    sc, poses, pm, srcs, need = scene_for(a, ref_views, dev, extra_views=[check_view] if (world > 1 and rank == 0) else [])
    gray = gray_f32(sc)
    d_gray = {i: eng.to_device(gray[i]) for i in need}
    my_bgr = list(ref_views) + ([check_view] if (world > 1 and rank == 0) else [])
    d_bgr = {r: eng.to_device(sc.images[r]['image'], torch.uint8) for r in my_bgr}
    Kinv32, Kinv64 = kinv32(sc.K), np.linalg.inv(sc.K)

    def cams_of(r, ss):
        return make_cams(sc.K, Kinv32, poses[r].R, poses[r].t, [poses[i].R for i in ss], [poses[i].t for i in ss])
    cams = {r: cams_of(r, srcs[r]) for r in my_bgr}
    dmin, dmax = sc.depth_range
    outs = {r: dict(depth=torch.empty((H, W), device=dev), normal=torch.empty((H, W, 3), device=dev),
                    confidence=torch.empty((H, W), device=dev)) for r in my_bgr}
    NS = max(1, min(a.streams, V))
    pts_buf = {r: torch.empty((H * W, 3), dtype=torch.float64, device=dev) for r in my_bgr}
    col_buf = {r: torch.empty((H * W, 3), dtype=torch.uint8, device=dev) for r in my_bgr}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    geometric = a.consistency == 'geometric'
    fp32_peak = eng.measure_fp32_peak()                                    # T lane-FMA/s, measured now, on this GPU

    def pm_job(r, arith):
        def job(slot):
            return eng.pm_view(d_gray[r], [d_gray[i] for i in srcs[r]], cams[r], a.patch, a.iters, a.samples, dmin, dmax,
                               seed=1234, view_id=r, out=outs[r], slot=slot, arith=arith)
        return job

    def fuse(r, slot=0):
        o = outs[r]
        return eng.mask_backproject(o['depth'], o['confidence'], d_bgr[r], Kinv64, poses[r].R, poses[r].t, 3.0, 0,
                                    points=pts_buf[r], colors=col_buf[r], sync=False, slot=slot)

    def view_job(r, arith):
        pj = pm_job(r, arith)

        def job(slot):
            pj(slot)
            return fuse(r, slot)
        return job

    state = {}

    def exchange_and_filter(time_it=False):
        """BASELINE configs[3] extension: all-gather of the depth AND normal maps of every view, then the geometric
        filter on this rank's views against the gathered depth maps of their sources (sources outside the computed set
        -- N < 8 computes only part of the turntable -- are skipped)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        all_depth = bdist.all_gather_maps(all_views, {r: outs[r]['depth'] for r in ref_views}, (H, W), torch.device(dev))
        all_normal = bdist.all_gather_maps(all_views, {r: outs[r]['normal'] for r in ref_views}, (H, W, 3), torch.device(dev))
        ev[1].record()
        checked = 0
        for r in ref_views:
            ss = [s for s in srcs[r] if s in pos_of]
            if not ss:
                continue
            geo = eng.geo_consistency(outs[r]['depth'], [all_depth[s] for s in ss], cams_of(r, ss), 1.0, 0.01)
            if geometric:
                outs[r]['confidence'].mul_((geo >= 2.0).to(torch.float32))
            checked += len(ss)
        ev[2].record()
        state['normals_gathered'] = int(all_normal[all_views[0]].numel()) * len(all_views)
        return ev, checked

    big = {}
    cnt_d = torch.zeros(V, dtype=torch.int64, device=dev)
    dummy = torch.zeros(1, dtype=torch.int64, device=dev)
    dev_t = torch.device(dev)

    def count_job(j, r, arith):
        pj = pm_job(r, arith)

        def job(slot):
            o = pj(slot)
            return eng.mask_count(o['depth'], o['confidence'], 3.0, 0, cnt_d[j:j + 1], slot=slot)
        return job

    def step(arith=ARITH):
        if world == 1:
            if geometric:
                eng.run_views([pm_job(r, arith) for r in ref_views], NS)
                exchange_and_filter()
                return [fuse(r) for r in ref_views]
            return eng.run_views([view_job(r, arith) for r in ref_views], NS)
        # N > 1 -- the job's only exchange under reference semantics: the fused clouds of all ranks' views on rank 0 in
        # sorted-view order (mvs_patchmatch.py:544-570).  Masks are counted on the device, the counts exchanged once, and
        # then every rank back-projects each view STRAIGHT INTO rank 0's cloud at its global offset (peer memory over
        # NVLink: the back-projection kernel is the transfer); NCCL point-to-point into the final slices is the fallback.
        if geometric:
            eng.run_views([pm_job(r, arith) for r in ref_views], NS)
            exchange_and_filter()
            for j, r in enumerate(ref_views):
                eng.mask_count(outs[r]['depth'], outs[r]['confidence'], 3.0, 0, cnt_d[j:j + 1])
        else:
            eng.run_views([count_job(j, r, arith) for j, r in enumerate(ref_views)], NS)
        cnts = cnt_d.cpu().tolist()
        counts = bdist.exchange_counts(len(all_views), {pos_of[r]: m for r, m in zip(ref_views, cnts)}, dev_t)
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        total = int(offs[-1])
        dest = bdist.peer_cloud(dev_t, total)

        def bp(r, pts, cols):
            o = outs[r]
            eng.mask_backproject(o['depth'], o['confidence'], d_bgr[r], Kinv64, poses[r].R, poses[r].t, 3.0, 0,
                                 points=pts, colors=cols, count=dummy, sync=False)
        if dest is not None:
            for r, m in zip(ref_views, cnts):
                o = int(offs[pos_of[r]])
                if m:
                    bp(r, dest[0][o:o + m], dest[1][o:o + m])
            bdist.peer_fence(dev_t)
            state['gather'] = 'peer stores into rank 0 (symmetric memory over NVLink)'
            if rank == 0:
                state.update(points=dest[0], colors=dest[1], counts=counts, offs=offs)
        else:
            state['gather'] = 'NCCL point-to-point into the final slices'
            if rank == 0:
                if 'p' not in big or big['p'].shape[0] < total:
                    big['p'] = torch.empty((total + (total >> 3), 3), dtype=torch.float64, device=dev)
                    big['c'] = torch.empty((total + (total >> 3), 3), dtype=torch.uint8, device=dev)
                fp, fc = big['p'][:total], big['c'][:total]
                for r, m in zip(ref_views, cnts):                           # own views: straight into place
                    o = int(offs[pos_of[r]])
                    if m:
                        bp(r, fp[o:o + m], fc[o:o + m])
                bdist.gather_into(counts, {}, fp, fc)
                state.update(points=fp, colors=fc, counts=counts, offs=offs)
            else:
                for r, m in zip(ref_views, cnts):
                    if m:
                        bp(r, pts_buf[r][:m], col_buf[r][:m])
                bdist.gather_into(counts, {pos_of[r]: (pts_buf[r][:m], col_buf[r][:m]) for r, m in zip(ref_views, cnts)})
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for e0, e1 in ev:
            flush.zero_()                       # L2 flush between timed iterations (not timed)
            e0.record()
            fn()
            e1.record()
        barrier()
        return max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in ev))

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ms = timed(step, a.steps, 0)
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    value = V * world * a.steps / (ms / 1e3)

    # ---- N>1: is the cloud rank 0 holds the right one?  Rank 0 recomputes a view owned by rank 1 (the Philox stream is
    # keyed by (seed, view, evaluation, pixel), so any rank reproduces any view) and compares its fused slice with the
    # gathered slice bit for bit; the total must equal the all-reduced sum of the per-view counts.
    gather_check = None
    if world > 1:
        barrier()
        if rank == 0:
            r = check_view
            pm_job(r, ARITH)(0)
            if geometric:
                gather_check = {'ok': None, 'note': 'skipped with --consistency geometric (needs every rank\'s maps)'}
            else:
                p, c, cnt = fuse(r)
                m = int(cnt.item())
                o = int(state['offs'][pos_of[r]])
                own = ref_views[0]
                pm_job(own, ARITH)(0)
                p0, c0, cnt0 = fuse(own)
                m0, o0 = int(cnt0.item()), int(state['offs'][pos_of[own]])
                ok = (m == int(state['counts'][pos_of[r]]) and m0 == int(state['counts'][pos_of[own]])
                      and torch.equal(state['points'][o:o + m], p[:m]) and torch.equal(state['colors'][o:o + m], c[:m])
                      and torch.equal(state['points'][o0:o0 + m0], p0[:m0]) and torch.equal(state['colors'][o0:o0 + m0], c0[:m0])
                      and int(state['points'].shape[0]) == int(sum(state['counts'])))
                gather_check = {'ok': bool(ok), 'points': int(state['points'].shape[0]), 'views': len(all_views),
                                'checked_view': int(r), 'owner_rank': 1, 'slice_points': m, 'gather': state.get('gather'),
                                'how': 'rank 0 recomputed a view owned by rank 1 (same seed / view id) and compared its fused '
                                       'slice of the gathered cloud bit for bit (points f64 + colours); own slice and total '
                                       'count checked too'}
        barrier()

    # ---- per-kernel durations: one more step right after the timed region, views serialised on one stream
    # (with several views in flight the kernels of different views overlap and a per-launch time is not
    # attributable), CUDA events on the launching stream around every launch (b200d_profile_begin/_end).
    flush.zero_()
    eng.profile_begin(65536)
    eng.run_views([view_job(r, ARITH) for r in ref_views], 1)
    torch.cuda.synchronize()
    kms, ktag = eng.profile_end()
    off = 10000 if a.arith == 'fast' else 0
    names = {off + 400 + a.patch: 'ncc_kernel<REFINE>', off + 300 + a.patch: 'ncc_kernel<SHIFT>',
             off + 200 + a.patch: 'ncc_kernel<CONF>', a.patch: 'ncc_kernel<STATS>', off + 2020: 'sample_kernel<REFINE>',
             off + 2010: 'sample_kernel<SHIFT>', off + 2001: 'sample_kernel<MAP,conf>', 3000: 'pm_init_kernel',
             4000: 'fuse kernels (3)'}
    tot_k = float(kms.sum())
    kernels = {names.get(int(tg), str(int(tg))): {'launches_per_step': int((ktag == tg).sum()),
                                                   'avg_ms': float(kms[ktag == tg].mean()),
                                                   'share_of_step': float(kms[ktag == tg].sum() / tot_k)}
               for tg in np.unique(ktag)}
    dom_ms = kernels['ncc_kernel<REFINE>']['avg_ms']
    smp_ms = kernels['sample_kernel<REFINE>']['avg_ms']
    peaks = load_json(os.path.join(ROOT, 'MEASURED_PEAKS.json')) or {}
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks else '6650 GB/s (of fallback)'
    # algorithmic HBM bytes of one cost evaluation = images 4(1+S) + state 2*(4+12+4) per pixel.
    alg_bytes = (44 + 4 * a.src) * H * W
    ach = alg_bytes / (dom_ms * 1e-3) / 1e9
    evals = H * W * a.src                                     # (pixel, source) evaluations per launch
    fma_peak_nominal = 148 * 128 * peaks.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
    k = a.patch
    # SURVEY 8(d) lane-op models per (pixel, source), FMA counted once.  Fast model W_alg(k) = 40 + 6(k-1): the sampling
    # kernel owns reprojection 8 + validity 5 + bilinear 14 = 27 of it, the NCC kernel products 2 + finalise 12 +
    # 6(k-1) separable sums.  Parity model 40 + 3k^2: the 3k^2 per-tap FMAs bit-exactness forces, in the NCC kernel.
    W_ALG_NCC, W_ALG_SAMPLE, W_PAR_NCC = 14 + 6 * (k - 1), 27, 14 + 3 * k * k
    pair_s = (dom_ms + smp_ms) * 1e-3
    ach_alg = evals * W_ALG_NCC / (dom_ms * 1e-3) / 1e12                    # T lane-op/s, dominant kernel, 8(d) model
    ach_par = evals * W_PAR_NCC / (dom_ms * 1e-3) / 1e12
    traffic_file = load_json(os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')) or {}
    traffic = (traffic_file.get('kernels', {}).get('ncc_kernel<11,REFINE>', {}) or {}).get('dram_bytes_per_launch') \
        if (k == 11 and a.arith == 'parity' and (H, W) == (1200, 1600)) else None
    roofline = {
        'bound': 'fp32', 'kernel': f'ncc_kernel<{k},REFINE>' + (' (B200D_FAST twin)' if a.arith == 'fast' else ''),
        'unit': 'T lane-op/s (an FMA counts once; x2 = TFLOP/s)',
        'achieved': ach_alg, 'peak': fp32_peak, 'frac': ach_alg / fp32_peak,
        'frac_definition': 'SURVEY 8(d) headline model: algorithmic lane-ops of the fast formulation (W_alg: this kernel\'s '
                           f'share {W_ALG_NCC} of 40 + 6(k-1) = {40 + 6 * (k - 1)} per pixel x source) / measured launch time / '
                           'measured FP32 peak',
        'achieved_parity': ach_par, 'frac_parity': ach_par / fp32_peak,
        'frac_parity_definition': f'the arithmetic bit-exactness forces: {W_PAR_NCC} lane-ops per pixel x source (3k^2 per-tap '
                                  'FMAs + 14) / the same time / the same peak',
        'pair': {'kernels': 'sample_kernel<REFINE> + ncc_kernel<REFINE> = one cost evaluation',
                 'ms': pair_s * 1e3,
                 'frac': evals * (40 + 6 * (k - 1)) / pair_s / 1e12 / fp32_peak,
                 'frac_parity': evals * (40 + 3 * k * k) / pair_s / 1e12 / fp32_peak},
        'whole_step': {'views_per_s': value / world, 'roofline_views_per_s_8d': 1.0 / max(
            (44 + 4 * a.src) * H * W * (a.iters * (2 + a.samples) + 1) / (hbm_peak * 1e9),
            (40 + 6 * (k - 1)) * evals * (a.iters * (2 + a.samples) + 1) / (fp32_peak * 1e12))},
        'peak_source': 'b200d_measure_fp32_peak: FFMA2 loop timed in this run on this GPU (of measured); nominal '
                       f'{fma_peak_nominal:.1f}',
        'traffic': traffic,
        'traffic_source': 'profiles/r02_ncu_traffic.json (profiles/summarize_ncu.py traffic <report>: dram__bytes_read.sum + '
                          'dram__bytes_write.sum per launch of one ncu --set full capture)' if traffic else
                          'null: no ncu capture of this exact kernel / shape on file',
        'launch_ms': dom_ms, 'launches_timed': kernels['ncc_kernel<REFINE>']['launches_per_step'],
        'kernel_share_of_step': kernels['ncc_kernel<REFINE>']['share_of_step'],
        'how': 'CUDA events on the launching stream around every launch, one serialised step right after the timed region',
        'note': 'contract lists hbm|tensor; this path is neither: no dense contraction (no tensor cores) and ~10 % of HBM '
                'peak -- the FP32 pipe binds it, so both fractions are quoted against the measured FP32 peak',
        'hbm': {'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                'algorithmic_bytes_per_launch': alg_bytes, 'peak_source': peak_src},
        'kernels': kernels}
    roofline['whole_step']['frac'] = roofline['whole_step']['views_per_s'] / roofline['whole_step']['roofline_views_per_s_8d']

    # ---- exchange leg (BASELINE configs[3] as written): all-gather of depth + normal maps, geometric filter
    exchange = None
    if not a.no_extras:
        for _ in range(2):
            exchange_and_filter()
        barrier()
        reps, t_ag, t_geo, checked = 3, 0.0, 0.0, 0
        for _ in range(reps):
            ev, checked = exchange_and_filter()
            barrier()
            t_ag += max_over_ranks(ev[0].elapsed_time(ev[1]))
            t_geo += max_over_ranks(ev[1].elapsed_time(ev[2]))
        recv = len(all_views) * H * W * 16                       # depth 4 B + normal 12 B per pixel, every view, per rank
        exchange = {'all_gather_ms': t_ag / reps, 'bytes_received_per_rank': recv,
                    'all_gather_gbs_per_rank': recv / (t_ag / reps * 1e-3) / 1e9 if world > 1 else None,
                    'maps': 'depth (H,W) + normal (H,W,3) fp32 of all %d views on every rank' % len(all_views),
                    'collective': 'ncclAllGather via torch.distributed.all_gather_into_tensor' if world > 1
                                  else 'single rank: local copy only',
                    'geo_filter_ms_per_rank': t_geo / reps, 'geo_sources_checked_per_rank': checked,
                    'in_timed_step': geometric,
                    'note': 'EXTENSION (north_star (c) / configs[3]); the reference\'s own consistency check is photometric and '
                            'needs no exchange.  Normals are consumed by schedule=\'viewprop\', depth by the geometric filter.'}

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
        dt = max_over_ranks(time.perf_counter() - t0)
        n_img = len(set(ref_views) | {s for r in ref_views for s in srcs[r]})
        e2e = {'value': V * world * a.steps / dt, 'unit': UNIT,
               'h2d_bytes_per_step': n_img * H * W * 4, 'd2h_bytes_per_step': V * 5 * H * W * 4,
               'api': 'core.mvs_patchmatch.PatchMatchMVS.compute_depth_maps (batch form of _patchmatch_cuda: pinned host '
                      'grays in, every distinct image uploaded once per step, host depth/normal/confidence out)'}

    # ---- end to end through reconstruct(): host photographs -> depth maps -> fusion -> gather -> post-filter -> host cloud
    e2e_rec = None
    if not a.no_extras and not a.no_e2e:
        e2e_rec = guarded(world, 'e2e_reconstruct', lambda: reconstruct_leg(a, dev, world, barrier, max_over_ranks))

    # ---- secondary metric: plane sweep Gpix*planes/s (BASELINE configs[2] shape)
    sweep = None
    if not a.no_plane_sweep:
        sweep = guarded(world, 'plane_sweep',
                        lambda: sweep_leg(a, eng, dev, rank, world, barrier, max_over_ranks, fp32_peak))

    # ---- the opt-in fast arithmetic beside the parity line (never the default; agreement measured by the tests)
    fast_mode = None
    if not a.no_extras and a.arith == 'parity' and not geometric:
        t_fast = timed(lambda: step(_cabi.FAST), 2, 2)
        v_fast = V * world * 2 / (t_fast / 1e3)
        step(ARITH)                                              # leave parity results in the buffers
        roof = roofline['whole_step']['roofline_views_per_s_8d']
        fast_mode = {'value': v_fast, 'unit': UNIT, 'steps': 2, 'arith': 'B200D_FAST (opt-in, NOT bit-exact, not the graded path)',
                     'frac_of_8d_roofline': v_fast / world / roof, 'north_star_target_views_per_s_per_gpu': 0.5 * roof,
                     'agreement_vs_oracle': load_json(os.path.join(ROOT, 'profiles', 'r02_fast_agreement.json')),
                     'agreement_source': 'profiles/r02_fast_agreement.json, written by tests/test_gpu_fast.py on a B200'}

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
                         f'through oracle/dense_oracle.c (OpenMP, {threads} threads), {dt:.1f} s',
               'reference_pytorch_cpu': REFERENCE_PYTORCH_CPU}

    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': config_dict(a, world),
            'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
            'gather_check': gather_check, 'exchange': exchange, 'e2e_reconstruct': e2e_rec, 'plane_sweep': sweep,
            'fast_mode': fast_mode}))
    if world > 1:
        dist.destroy_process_group()


def reconstruct_leg(a, dev, world, barrier, max_over_ranks, n_views=16):
    """PatchMatchMVS.reconstruct() on the first `n_views` views of the turntable: host BGR photographs in (uploaded and
    prepared on the device), depth maps, masks, fusion, NCCL gather to rank 0, post-filter, host cloud out."""
    import torch
    from b200dense.scenes import make_scene
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    sc = make_scene('turntable', N_SCENE_VIEWS, a.width, a.height, arc_deg=360.0 / N_SCENE_VIEWS, seed=0, device=dev,
                    render=list(range(n_views)))
    images = [{'image': sc.images[i]['image']} for i in range(n_views)]
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in range(n_views)}
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    th = np.random.default_rng(0).normal(size=(400, 3))
    sparse = 1.5 * th / np.linalg.norm(th, axis=1, keepdims=True)           # "SfM points" on the object
    with quiet():
        pm = PatchMatchMVS(cam, scale=1.0, patch_size=a.patch, num_iterations=a.iters, num_samples=a.samples, min_views=3,
                           num_src_views=a.src, seed=1234, arith=a.arith, device=dev)
        pts, _ = pm.reconstruct(images, poses, sparse_points=sparse)        # warm-up
        barrier()
        t0 = time.perf_counter()
        pts, cols = pm.reconstruct(images, poses, sparse_points=sparse)
        barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    tm = dict(pm.last_timings)
    # the same through the view-propagation extension: all views in lock step, depth + normal maps all-gathered after
    # every iteration (NCCL at N>1), one proposal evaluation per source view
    with quiet():
        pv = PatchMatchMVS(cam, scale=1.0, patch_size=a.patch, num_iterations=a.iters, num_samples=a.samples, min_views=3,
                           num_src_views=a.src, seed=1234, arith=a.arith, device=dev, schedule='viewprop')
        pv.reconstruct(images, poses, sparse_points=sparse)
        barrier()
        t0 = time.perf_counter()
        vpts, _ = pv.reconstruct(images, poses, sparse_points=sparse)
        barrier()
    dtv = max_over_ranks(time.perf_counter() - t0)
    viewprop = {'value': n_views / dtv, 'unit': UNIT, 'seconds': dtv, 'points_out': int(len(vpts)),
                'evaluations_per_view': a.iters * (2 + a.samples + a.src) + 1,
                'exchange': f'{a.iters} all-gathers of depth + normal maps of all {n_views} views',
                'note': 'EXTENSION schedule=\'viewprop\' (no dackey-wav counterpart); bit-equal to its NumPy-on-oracle definition '
                        '(tests/test_gpu_viewprop.py), 2-rank result == 1-rank result (tests/test_gpu_multi.py)'}
    return {'value': n_views / dt, 'unit': UNIT, 'views': n_views, 'seconds': dt, 'viewprop_schedule': viewprop,
            'h2d_bytes_per_step': n_views * a.height * a.width * 3, 'd2h_bytes_per_step': int(len(pts)) * 27,
            'points_out': int(len(pts)), 'phases_rank0_s': tm,
            'api': 'core.mvs_patchmatch.PatchMatchMVS.reconstruct(images, poses, sparse_points): host BGR photographs in, '
                   'fused + filtered host cloud out (views sharded over the ranks, clouds gathered to rank 0)'}


def sweep_leg(a, eng, dev, rank, world, barrier, max_over_ranks, fp32_peak):
    import torch
    from b200dense.engine import kinv32, make_cams
    from b200dense.scenes import gray_f32, make_scene
    from core.camera import Camera, CameraPose
    from core.dense_stereo import DenseStereoReconstructor
    Ws, Hs, D, N, k = 1920, 1080, 256, 6, 7
    ARITH = 1 if a.arith == 'fast' else 0
    nb_views = list(range(0, 10))
    scs = make_scene('turntable', N_SCENE_VIEWS, Ws, Hs, arc_deg=360.0 / N_SCENE_VIEWS, seed=0, device=dev, render=nb_views)
    gs = gray_f32(scs)
    dg = {i: eng.to_device(gs[i]) for i in nb_views}
    refs_s = [3, 4, 5, 6]                                  # 4 reference views per rep, 2 in flight
    depths = (1.0 / np.linspace(1 / scs.depth_range[1], 1 / scs.depth_range[0], D)).astype(np.float32)
    Ki = kinv32(scs.K)

    def nbs_of(r):
        return [r - 1, r + 1, r - 2, r + 2, r - 3, r + 3]  # 6 nearest cameras on the circle

    def sweep_job(r):
        nbs = nbs_of(r)
        cs = make_cams(scs.K, Ki, scs.poses[r][0], scs.poses[r][1], [scs.poses[i][0] for i in nbs],
                       [scs.poses[i][1] for i in nbs])
        return lambda slot: eng.plane_sweep(dg[r], [dg[i] for i in nbs], cs, depths, k, 0.8, slot=slot, arith=ARITH)
    sjobs = [sweep_job(r) for r in refs_s]
    eng.run_views(sjobs, a.sweep_streams)                  # warm-up: every view once (kernels loaded, scratch slots allocated)
    barrier()
    eng.sweep_fallbacks(reset=True)
    reps, rep_ms = 3, []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        eng.run_views(sjobs, a.sweep_streams)
        e1.record()
        barrier()
        rep_ms.append(max_over_ranks(e0.elapsed_time(e1)))
    tms = float(np.median(rep_ms))
    n_fb = eng.sweep_fallbacks() / reps
    gpp = world * len(refs_s) * Hs * Ws * D / (tms / 1e3) / 1e9
    # per-kernel times of one serialised view
    eng.profile_begin(4096)
    eng.run_views(sjobs[:1], 1)
    torch.cuda.synchronize()
    kms, ktag = eng.profile_end()
    kn = {2032: 'sample_kernel<PLANE>', 12032: 'sample_fast_kernel<PLANE>', 20500 + k: 'ncc_sep_kernel<SWEEP,bounded>',
          500 + k: 'ncc_kernel<SWEEP,chains>', 10500 + k: 'ncc_sep_kernel<SWEEP,fast>', k: 'ncc_kernel<STATS>'}
    tot = float(kms.sum())
    kernels = {kn.get(int(t), str(int(t))): {'launches_per_view': int((ktag == t).sum()), 'avg_ms': float(kms[ktag == t].mean()),
                                             'p50_ms': float(np.median(kms[ktag == t])), 'max_ms': float(kms[ktag == t].max()),
                                             'share_of_view': float(kms[ktag == t].sum() / tot)} for t in np.unique(ktag)}
    evals_per_s = gpp * 1e9 * N / world
    sweep = {'metric': 'plane_sweep_gpix_planes_per_s', 'value': gpp, 'unit': 'Gpix*planes/s',
             'config': f'{Ws}x{Hs}, D={D}, N={N}, {k}x{k} NCC, {len(refs_s)} ref views per GPU ({a.sweep_streams} in flight), '
                       f'{a.arith} arithmetic (BASELINE configs[2] shape)',
             'ms_per_view': tms / len(refs_s), 'evals_per_s': gpp * 1e9 * N, 'rep_ms': rep_ms,
             'timing': f'median of {reps} repetitions of {len(refs_s)} views (CUDA events, max over ranks), after one warm-up pass',
             'roofline': {'bound': 'fp32', 'unit': 'T lane-op/s', 'peak': fp32_peak,
                          'achieved': evals_per_s * (40 + 6 * (k - 1)) / 1e12,
                          'frac': evals_per_s * (40 + 6 * (k - 1)) / 1e12 / fp32_peak,
                          'frac_definition': 'SURVEY 8(d): W_alg(7) = 76 lane-ops per pixel x plane x neighbour; roofline '
                                             f'{fp32_peak * 1e12 / (76 * N) / 1e9:.1f} Gpix*planes/s per GPU at N=6',
                          'frac_parity': evals_per_s * (40 + 3 * k * k) / 1e12 / fp32_peak,
                          'frac_parity_definition': 'the chains-everywhere count 40 + 3k^2 = 187; the bounded decision '
                                                    'evaluates the chains only where the bound is inconclusive, so this '
                                                    'fraction may exceed what chains-everywhere could reach',
                          'kernels': kernels},
             'bounded_decision': {'exact_redo_fraction_of_outputs':
                                  n_fb / float(Ws * Hs * N * D * len(refs_s)) if a.arith == 'parity' else None,
                                  'note': 'B200D_PARITY sweep: `ncc > thresh` decided from separable sums + rigorous error '
                                          'bound, undecided outputs redo the exact chains; counts / arg-max bit-identical '
                                          '(tests/test_gpu_sweep_hybrid.py)'}}
    # e2e through the reference-facing seam: host grays in, host depth / confidence out
    cam = Camera(K=scs.K.copy(), dist=np.zeros(5))
    with quiet():
        ds = DenseStereoReconstructor(cam, scale=1.0, num_depths=D, patch_size=k, device=dev, arith=a.arith)
    pinned = {i: torch.from_numpy(gs[i]).pin_memory() for i in nb_views}
    processed = {i: {'gray': pinned[i].numpy(), 'color': None, 'shape': (Hs, Ws)} for i in nb_views}
    poses = {i: CameraPose(R=scs.poses[i][0], t=scs.poses[i][1]) for i in scs.poses}

    nbr_of = {r: nbs_of(r) for r in refs_s}
    ds.views_in_flight = a.sweep_streams

    def e2e_step():
        maps = ds.compute_depth_maps(processed, poses, refs_s, nbr_of, scs.depth_range[0], scs.depth_range[1])
        assert len(maps) == len(refs_s) and maps[refs_s[0]][0].shape == (Hs, Ws)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        e2e_step()
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    n_img = len({i for r in refs_s for i in [r] + nbr_of[r]})
    sweep['e2e'] = {'value': world * 2 * len(refs_s) * Hs * Ws * D / dt / 1e9, 'unit': 'Gpix*planes/s',
                    'h2d_bytes_per_step': n_img * Hs * Ws * 4, 'd2h_bytes_per_step': len(refs_s) * 2 * Hs * Ws * 4,
                    'api': 'core.dense_stereo.DenseStereoReconstructor.compute_depth_maps (batch form of _compute_depth_map_gpu: '
                           'pinned host grays in, every distinct image uploaded once per step, host depth map + confidence out; '
                           f'{len(refs_s)} views per step, {a.sweep_streams} in flight)'}
    # CPU arm of this metric: BASELINE configs[0] in full through the port (5 views 640x480, 64 planes, 7x7, N=4)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import oracle as orc
        threads = oracle_threads()
        s0 = make_scene('plane', 5, 640, 480, focal=600.0, arc_deg=8.0, seed=0)
        g0 = gray_f32(s0)
        d0 = (1.0 / np.linspace(1 / 20.0, 1 / 0.5, 64)).astype(np.float32)
        K0 = kinv32(s0.K)
        t0 = time.perf_counter()
        for r in range(5):
            nb = [i for i in sorted(range(5), key=lambda j: (abs(j - r), j)) if i != r][:4]
            oc = orc.Cams(s0.K, K0, s0.poses[r][0], s0.poses[r][1], [s0.poses[i][0] for i in nb], [s0.poses[i][1] for i in nb])
            orc.plane_sweep(g0[r], [g0[i] for i in nb], oc, d0, 7, 0.8)
        dt = time.perf_counter() - t0
        sweep['cpu_baseline'] = {'value': 5 * 640 * 480 * 64 / dt / 1e9, 'unit': 'Gpix*planes/s', 'cores': threads, 'kind': 'port',
                                 'sample': f'BASELINE configs[0] in full: 5 views 640x480, 64 planes, 7x7 NCC, N=4 through '
                                           f'oracle/dense_oracle.c ({threads} threads), {dt:.1f} s',
                                 'reference_pytorch_cpu_gpix_planes_per_s': REFERENCE_PYTORCH_CPU['plane_sweep_gpix_planes_per_s_cfg0']}
    return sweep


def run_full4k(a):
    """BASELINE configs[4]: the full dense stage on an N-view 3840x2160 turntable -- reconstruct() (upload + prepare,
    PatchMatch, masks, fusion, gather) -> post-filter -> save_ply -- with the share of each phase."""
    import tempfile
    import torch
    import torch.distributed as dist
    from b200dense.scenes import make_scene
    from core.camera import Camera, CameraPose
    from core.mvs_patchmatch import PatchMatchMVS
    from core.utils import save_ply
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    init_dist(world, dev)
    Wf, Hf, nv = 3840, 2160, a.full4k_views
    t0 = time.perf_counter()
    sc = make_scene('turntable', nv, Wf, Hf, arc_deg=360.0 / nv, seed=0, device=dev, render=list(range(nv)))
    images = [{'image': sc.images[i]['image']} for i in range(nv)]
    sc.depth_gt = []
    poses = {i: CameraPose(R=sc.poses[i][0], t=sc.poses[i][1]) for i in range(nv)}
    cam = Camera(K=sc.K.copy(), dist=np.zeros(5))
    th = np.random.default_rng(0).normal(size=(2000, 3))
    sparse = 1.5 * th / np.linalg.norm(th, axis=1, keepdims=True)
    t_scene = time.perf_counter() - t0
    torch.cuda.empty_cache()
    with quiet():
        pm = PatchMatchMVS(cam, scale=1.0, patch_size=a.patch, num_iterations=a.iters, num_samples=a.samples,
                           min_views=a.full4k_min_views, num_src_views=a.src, seed=1234, arith=a.arith, device=dev)
        # warm-up on a 6-view subset (not timed): kernels loaded, scratch and staging buffers allocated, sort / select
        # workspaces of the post-filter created -- what a second reconstruction in the same process would find
        wn = min(6 * world, nv)
        wp, wc = pm.reconstruct(images[:wn], {i: poses[i] for i in range(wn)}, sparse_points=sparse)
        if rank == 0 and len(wp):
            save_ply(wp, wc, os.path.join(tempfile.mkdtemp(prefix='b200dense_'), 'warm.ply'))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with quiet():
        pts, cols = pm.reconstruct(images, poses, sparse_points=sparse)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_rec = time.perf_counter() - t0
    tm = dict(pm.last_timings)
    t_ply, ply_bytes = 0.0, 0
    if rank == 0:
        out = a.ply_out or os.path.join(tempfile.mkdtemp(prefix='b200dense_'), 'dense_mvs.ply')
        t0 = time.perf_counter()
        with quiet():
            save_ply(pts, cols, out)
        t_ply = time.perf_counter() - t0
        ply_bytes = os.path.getsize(out)
        if not a.ply_out:
            os.remove(out)
    total = t_rec + t_ply
    if rank == 0:
        phases = {'depth_maps_incl_upload_prepare': tm.get('depth_maps_s'), 'fusion_and_gather': tm.get('fusion_gather_s'),
                  'post_filter_and_download': tm.get('filter_download_s'), 'save_ply': t_ply}
        print(json.dumps({
            'metric': 'full_dense_stage_views_per_s_3840x2160', 'value': nv / total, 'unit': UNIT, 'n_gpus': world,
            'steps': 1, 'warmup': 0, 'ms_per_step': total * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'BASELINE configs[4]: {nv}-view synthetic turntable at {Wf}x{Hf}, PatchMatch S={a.src}, '
                                   f'{a.patch}x{a.patch}, {a.iters} iters x (2 + {a.samples}) + confidence, conf>={a.full4k_min_views} fusion, '
                                   f'gather to rank 0, 95th-percentile + 0.01-voxel filter, ASCII PLY',
                       'arithmetic': a.arith, 'parallelism': f'view-sharded dp{world}'},
            'seconds': {'reconstruct': t_rec, 'save_ply': t_ply, 'total': total, 'scene_synthesis_not_timed': t_scene},
            'host_side_seconds_inside_depth_maps': {k2: tm.get(k2) for k2 in ('host_select_s', 'host_stage_upload_s', 'host_enqueue_s')},
            'phase_seconds': phases, 'phase_share': {k2: (v / total if v is not None else None) for k2, v in phases.items()},
            'raw_points': tm.get('raw_points'), 'points_out': int(len(pts)), 'ply_bytes': ply_bytes,
            'e2e': {'value': nv / total, 'unit': UNIT, 'h2d_bytes_per_step': nv * Hf * Wf * 3,
                    'd2h_bytes_per_step': int(len(pts)) * 27}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
