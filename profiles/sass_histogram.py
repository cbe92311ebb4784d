#!/usr/bin/env python
"""SASS opcode histogram of the shipped library (evidence that the hot kernels are sm_100a code with TMA + FFMA2).

    python profiles/sass_histogram.py [path/to/libb200dense.so] > profiles/r02_sass_opcodes.txt

Runs `cuobjdump -sass` (works without a GPU) and prints, per kernel family, the instruction count of the opcodes that
matter here: UTMALDG (TMA tensor loads), SYNCS (mbarrier), FFMA2 / FMUL2 / FADD2 (packed fp32), FFMA, MUFU, LDS / STS,
LDG / STG, plus the cubin's architecture line.  No tensor-core opcodes are expected: nothing on this path is a
contraction."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, '3d-reconstruction-tool_b200', 'csrc', 'libb200dense.so')
WATCH = ['UTMALDG', 'SYNCS', 'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'FMUL', 'FADD', 'MUFU', 'LDS', 'STS', 'LDG', 'STG', 'SHFL',
         'ATOMG', 'RED', 'BAR', 'HMMA', 'UTCMMA', 'UTCHMMA', 'IMAD', 'LOP3', 'ISETP', 'FSETP']


def family(name):
    for key, fam in (('ncc_sep_kernel', 'ncc_sep_kernel (separable sums: FAST mode / bounded plane sweep)'),
                     ('ncc_kernel', 'ncc_kernel (exact chains)'), ('sample_fast_kernel', 'sample_fast_kernel'),
                     ('sample_kernel', 'sample_kernel'), ('fuse_', 'fuse kernels'), ('geo_consistency', 'geo_consistency_kernel'),
                     ('viewprop_cand', 'viewprop_cand_kernel'), ('knn_mean', 'knn_mean_dist_kernel'),
                     ('prepare_image', 'prepare_image_kernel'), ('pm_init', 'pm_init_kernel'), ('fma_peak', 'fma_peak_kernel'),
                     ('selftest', 'selftest kernels')):
        if key in name:
            return fam
    return 'other'


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r'arch = (sm_\w+)', out)))
    per = collections.defaultdict(collections.Counter)
    nk = collections.Counter()
    fam = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fam = family(m.group(1))
            nk[fam] += 1
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and fam:
            per[fam][m.group(1)] += 1
    print(f'library: {os.path.relpath(LIB, ROOT)}')
    print(f'cubin architectures: {", ".join(arch)}')
    for fam in sorted(per, key=lambda f: -sum(per[f].values())):
        c = per[fam]
        tot = sum(c.values())
        print(f'\n== {fam}: {nk[fam]} kernels, {tot} SASS instructions')
        print('   ' + '  '.join(f'{op}={c[op]}' for op in WATCH if c[op]))
        top = ', '.join(f'{op} {n}' for op, n in c.most_common(8))
        print(f'   top: {top}')
    tensor = sum(per[f][op] for f in per for op in ('HMMA', 'UTCMMA', 'UTCHMMA'))
    print(f'\ntensor-core instructions in the library: {tensor} (expected 0: no dense contraction on this path)')


if __name__ == '__main__':
    main()
