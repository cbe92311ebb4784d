"""b200dense -- B200-native engine for the dense-reconstruction hot path (plane sweep,
PatchMatch MVS, consistency mask + back-projection) of dackey-wav/3d-reconstruction-tool.

Layout
  _cabi.py   ctypes binding of csrc/libb200dense.so (include/b200dense.h)
  engine.py  DenseEngine: device buffers / streams via torch, compute via the C ABI
  scenes.py  synthetic scenes with known poses (tests, fixtures, benchmarks)

The reference-facing drop-in classes live in ../src/core (same module names as the reference).
"""
from ._cabi import B200DenseError, FAST, PARITY, PARITY_CHAINS, PARITY_CHECK, LIB_PATH  # noqa: F401
