#!/usr/bin/env python
"""Build libb200dense.so in-tree with nvcc for sm_100a (no torch involved).

    python 3d-reconstruction-tool_b200/csrc/build.py [--force] [--verbose]

-fmad=false: no implicit contraction; FMAs are issued only where the source calls __fmaf_rn,
which is what makes the parity arithmetic bit-exact.  No -use_fast_math.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, 'libb200dense.so')
SRCS = ['b200dense.cu']
DEPS = SRCS + ['device_math.cuh', 'sample_kernels.cuh', 'ncc_kernels.cuh', 'fast_kernels.cuh', 'diag_kernels.cuh', 'fuse_kernels.cuh', 'geo_kernels.cuh', 'prep_kernels.cuh', 'knn_kernels.cuh',
               os.path.join(ROOT, 'include', 'b200dense.h')]


def nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in DEPS)


def build(force=False, verbose=False, out=None, defines=()):
    """Default build -> csrc/libb200dense.so.  `out` + `defines` build a tuning variant elsewhere."""
    if out is None and not force and not stale():
        return OUT
    if out is not None:
        OUT_ = out
    else:
        OUT_ = OUT
    cmd = [nvcc(), *[f'-D{d}' for d in defines], '-shared', '-Xcompiler', '-fPIC,-fvisibility=hidden,-fopenmp', '-std=c++17', '-O3', '-lineinfo',
           '-gencode', 'arch=compute_100a,code=sm_100a', '-fmad=false', '-ccbin', '/usr/bin/g++',
           '-I', os.path.join(ROOT, 'include'), '-I', HERE, '-o', OUT_] + [os.path.join(HERE, s) for s in SRCS] + ['-lgomp']
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    subprocess.check_call(cmd)
    return OUT_


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
