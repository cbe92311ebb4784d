# compare builds of the library on the fast-mode PatchMatch step (B200DENSE_LIB selects a variant .so)
for v in base pm1; do
  if [ "$v" = base ]; then unset B200DENSE_LIB; else export B200DENSE_LIB=/root/repo/profiles/microbench/variants/$v.so; fi
  for st in 4 6; do
  python bench.py --arith fast --no-cpu-baseline --no-extras --no-e2e --no-plane-sweep --steps 3 --streams $st > gpurun_out/varf_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/varf_$v.json')); print('$v fast streams $st', round(d['value'],1), {k:round(x['avg_ms'],4) for k,x in d['roofline']['kernels'].items() if 'REFINE' in k})"
  done
done
unset B200DENSE_LIB
python bench.py --no-cpu-baseline --no-extras --no-e2e --steps 3 > gpurun_out/var_base.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/var_base.json')); s=d['plane_sweep']; print('base parity', round(d['value'],1), round(s['value'],2))"
python -m pytest tests/test_gpu_fast.py tests/test_gpu_sweep_hybrid.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
