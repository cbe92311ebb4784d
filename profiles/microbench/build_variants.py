#!/usr/bin/env python
"""Build tuning variants of the CUDA library into profiles/microbench/variants/ (git-ignored .so files)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location('b', os.path.join(ROOT, '3d-reconstruction-tool_b200', 'csrc', 'build.py'))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
OUT = os.path.join(ROOT, 'profiles', 'microbench', 'variants')
os.makedirs(OUT, exist_ok=True)
for name, defs in [a.split(':') for a in sys.argv[1:]]:
    d = [x for x in defs.split(',') if x]
    print(name, d)
    b.build(force=True, verbose=False, out=os.path.join(OUT, f'{name}.so'), defines=d)
