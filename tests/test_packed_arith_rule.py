"""CPU: the packed fp32x2 path of the parity kernels must never feed a packed PRODUCT straight into a packed add / sub.

ptxas 12.9 contracts `mul.rn.f32x2` followed by `add.rn.f32x2` / `sub.rn.f32x2` into one FFMA2 despite the explicit .rn
and -fmad=false (profiles/microbench/f32x2_contract.cu) -- one rounding instead of two, i.e. not the
reference's arithmetic.  The rule is enforced here on the source of project_sample_fast2 (device_math.cuh), and the
instruction counts of the built library are checked against the counts the source implies (a contraction would turn an
FMUL2 + FADD2 pair into an extra FFMA2)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, '3d-reconstruction-tool_b200', 'csrc')


def _body():
    src = open(os.path.join(CSRC, 'device_math.cuh')).read()
    start = src.index('__device__ __forceinline__ void project_sample_fast2(')
    end = src.index('// ---- Philox4x32-10', start)
    return src[start:end]


def test_no_packed_product_is_an_operand_of_a_packed_add_or_sub():
    body = _body()
    # names bound to the result of f2mul(...)
    products = set(re.findall(r'(?:const\s+)?f32x2\s+(\w+)\s*=\s*f2mul\(', body))
    products |= set(re.findall(r',\s*(\w+)\s*=\s*f2mul\(', body))
    products |= set(re.findall(r'^\s*(\w+)\s*=\s*f2mul\(', body, flags=re.M))
    assert {'px0', 'py0', 'U2', 'V2', 'gq0', 'hq0', 'nw', 'ne', 'sw', 'se', 'a', 'o'} <= products
    # direct operands of every f2add / f2sub: never a product (neither a nested f2mul(...) nor a variable bound to one);
    # products nested deeper, inside an f2fma operand, are multiplicands / addends of that FMA and cannot be contracted.
    # `a` and `o` are reassigned by f2fma before they reach an add, so they are checked by position below
    n_calls = 0
    for m in re.finditer(r'\bf2(?:add|sub)\(', body):
        depth, args, cur = 1, [], ''
        for ch in body[m.end():]:
            if ch == '(':
                depth += 1
            elif ch == ')':
                depth -= 1
                if depth == 0:
                    args.append(cur.strip())
                    break
            if ch == ',' and depth == 1:
                args.append(cur.strip())
                cur = ''
            else:
                cur += ch
        assert len(args) == 2, args
        n_calls += 1
        for arg in args:
            assert not arg.startswith('f2mul('), arg
            assert arg not in (products - {'a', 'o'}), (arg, m.group(0))
    assert n_calls >= 14
    # the accumulators: between `a = f2mul(` / `o = f2mul(` and the next f2add / f2sub that reads them there is an f2fma
    for var in ('a', 'o'):
        for m in re.finditer(r'\b%s\s*=\s*f2mul\(' % var, body):
            rest = body[m.end():]
            use = re.search(r'f2(?:add|sub)\(\s*%s\b' % var, rest)
            if use:
                assert re.search(r'\b%s\s*=\s*f2fma\(' % var, rest[:use.start()]), var
    # the one product that IS subtracted from (ix - floor(ix)) is formed with scalar multiplies
    assert re.search(r'ix0\s*=\s*mul\(hx0,\s*wm1\)', body) and 'f2pack(ix0, ix1)' in body


@pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='cuobjdump not available')
def test_built_plane_sampling_kernel_has_the_packed_counts_the_source_implies():
    lib = os.path.join(CSRC, 'libb200dense.so')
    if not os.path.exists(lib):
        pytest.skip('library not built')
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    parts = re.split(r'\n\s*Function : ', sass)
    hit = [p for p in parts if p.startswith('_ZN5b200d13sample_kernelILi3ELi2ELi6EEE')]       # <C_PLANE, B_DEPTH, 6 sources>
    assert len(hit) == 1
    ops = re.findall(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', hit[0])
    # per pair of sources: 21 FFMA2, 18 FMUL2, 16 FADD2 (to_source 6/3/3, reciprocal 2/0/1, divisions 4/2/0, to_pixel 2/2/2,
    # grid round trip 4/6/4, bilinear 3/5/6); three pairs, nothing else in this kernel is packed
    assert (ops.count('FFMA2'), ops.count('FMUL2'), ops.count('FADD2')) == (63, 54, 48)
