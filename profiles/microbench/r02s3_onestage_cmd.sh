# one sample stage + 4 resident CTAs for the bounded sweep kernel (variant one4) vs the shipped 2-stage / 3-CTA kernel
V=$PWD/profiles/microbench/variants
for lib in default one4; do
  if [ $lib = default ]; then unset B200DENSE_LIB; else export B200DENSE_LIB=$V/$lib.so; fi
  timeout 300 python -m pytest tests/test_gpu_sweep_hybrid.py tests/test_gpu_public_api.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -1
  timeout 200 python bench.py --no-cpu-baseline --no-extras --no-e2e --steps 2 > gpurun_out/onestage_$lib.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/onestage_$lib.json')); ps=d['plane_sweep']; print('$lib', round(ps['value'],2), [round(x,1) for x in ps['rep_ms']], {k:round(v['avg_ms'],4) for k,v in ps['roofline']['kernels'].items()}, ps['bounded_decision']['exact_redo_fraction_of_outputs'])"
done
