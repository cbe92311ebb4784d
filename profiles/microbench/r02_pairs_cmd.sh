# Round-2 (session 3): packed two-source sampling (FFMA2/FMUL2/FADD2) vs the scalar kernel, and its residency.
# Variants built by: python profiles/microbench/build_variants.py base:B200D_SAMPLE_PAIRS=0 p6:B200D_SAMPLE_MIN_BLOCKS=6 p5:B200D_SAMPLE_MIN_BLOCKS=5 p4:B200D_SAMPLE_MIN_BLOCKS=4
V=profiles/microbench/variants
B200DENSE_LIB=$PWD/$V/p5.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_sweep_hybrid.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/microbench/time_kernels.py $V/base.so $V/p6.so $V/p5.so $V/p4.so 2>&1 | tail -12
for v in base p6 p5 p4; do
  B200DENSE_LIB=$PWD/$V/$v.so timeout 300 python bench.py --no-cpu-baseline --no-extras --no-e2e --steps 5 > gpurun_out/pairs_$v.json 2>gpurun_out/pairs_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/pairs_$v.json')); ps=d.get('plane_sweep') or {}
print('$v', round(d['value'],1), 'views/s; sweep', round(ps.get('value',0),2), {k:round(x['avg_ms'],4) for k,x in d['roofline']['kernels'].items() if 'sample' in k}, {k:round(x['avg_ms'],4) for k,x in (ps.get('roofline',{}).get('kernels',{})).items() if 'sample' in k})"
done
