"""CPU: the benchmark contract of the reference (CPU) arm -- one JSON line with the keys the driver reads, the same
`config` object the b200 arm prints, all host threads regardless of OMP_NUM_THREADS, `--warmup` honoured, ranks > 0 of a
torchrun launch silent.  (The b200 arm needs a GPU; its line is checked on the box by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ['--impl', 'reference', '--steps', '1', '--warmup', '1', '--width', '96', '--height', '72', '--gpus', '2']


def run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + ARGS, capture_output=True, text=True, env=env,
                          timeout=300)


def test_reference_arm_json_line():
    r = run({'OMP_NUM_THREADS': '1', 'RANK': '0'})                    # what torchrun exports
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['metric'] == 'patchmatch_views_per_s_96x72' and d['unit'] == 'views/s'
    assert d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 2 and d['higher_is_better'] is True
    assert d['vs_baseline'] is None and d['value'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and cb['sample']
    assert cb['cores'] == len(os.sched_getaffinity(0))                # not the launcher's OMP_NUM_THREADS=1
    assert d['e2e'] == {'value': d['value'], 'unit': 'views/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # same config object as the b200 arm
    sys.path.insert(0, ROOT)
    import bench
    a = bench.parse.__globals__['argparse'].Namespace(width=96, height=72, views_per_gpu=8, src=5, patch=11, iters=3,
                                                      samples=8, streams=4, arith='parity', consistency='photometric')
    assert d['config'] == bench.config_dict(a, 2)


def test_other_ranks_print_nothing():
    r = run({'RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''
