"""CPU: the C-ABI library loads and exports every symbol include/b200dense.h declares.
No compute call is made (no GPU here); argument validation paths that do not touch CUDA are checked."""
import ctypes as C
import os
import re

import pytest

from b200dense import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'b200dense.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200d_[a-z0-9_]+)\s*\(', src)))


def test_library_built():
    assert os.path.exists(_cabi.LIB_PATH), 'run python 3d-reconstruction-tool_b200/csrc/build.py'


def test_every_declared_symbol_is_exported_and_bound():
    lib = _cabi.load()
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/b200dense.h but not exported'
        assert s in _cabi.PROTOTYPES, f'{s} has no ctypes prototype'
    assert sorted(_cabi.PROTOTYPES) == syms


def test_abi_version_and_struct_layout():
    lib = _cabi.load()
    assert lib.b200d_abi_version() == _cabi.ABI_VERSION == 3
    # b200d_cams: 9+9+9+3 floats, int32, 8*9 + 8*3 floats
    assert C.sizeof(_cabi.Cams) == 4 * (30 + 1 + 72 + 24)
    assert C.sizeof(_cabi.PmParams) == 4 * 4 + 2 * 8 + 4 * 4
    assert lib.b200d_scratch_bytes(1200, 1600) >= 16 * 1200 * 1600 * 4
    assert lib.b200d_scratch_bytes(0, 10) == 0
    assert lib.b200d_fuse_scratch_bytes(10, 10) > 0


def test_argument_errors_do_not_touch_cuda():
    lib = _cabi.load()
    rc = lib.b200d_ref_stats(None, 10, 10, 4, None, None, None, None)     # even patch
    assert rc == -1 and b'patch' in lib.b200d_last_error()
    rc = lib.b200d_ref_stats(None, 10, 10, 5, None, None, None, None)     # null pointers
    assert rc == -1 and b'null' in lib.b200d_last_error()
    with pytest.raises(_cabi.B200DenseError):
        _cabi.check(rc)


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from b200dense.engine import DenseEngine
    with pytest.raises(_cabi.B200DenseError):
        DenseEngine('cuda:0')


def test_integration_md_structs_match_the_binding():
    """The ctypes structures INTEGRATION.md tells a reference maintainer to write are the ones the tested binding uses
    (same field names, order and C types), and every entry point its table names is exported."""
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    ns = {}
    block = re.search(r'class Cams\(C\.Structure\):.*?lib\.b200d_last_error\.restype', text, flags=re.S).group(0)
    block = block.rsplit('\n', 2)[0]
    exec('import ctypes as C\n' + block, ns)                      # noqa: S102  (our own documentation snippet)
    for name in ('Cams', 'PmParams', 'Rng'):
        doc, real = ns[name], getattr(_cabi, name)
        assert [(f[0], f[1]) for f in doc._fields_] == [(f[0], f[1]) for f in real._fields_], name
        assert C.sizeof(doc) == C.sizeof(real)
    lib = _cabi.load()
    types = {'b200d_status', 'b200d_cams', 'b200d_pm_params', 'b200d_rng', 'b200d_arith'}
    for sym in set(re.findall(r'`(b200d_[a-z0-9_]+)', text)) - types:
        assert hasattr(lib, sym), sym
