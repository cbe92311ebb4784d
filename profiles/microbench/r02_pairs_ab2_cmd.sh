V=profiles/microbench/variants
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do for v in base p6 p5; do
  B200DENSE_LIB=$PWD/$V/$v.so timeout 300 python bench.py --no-cpu-baseline --no-extras --no-e2e --no-plane-sweep --steps 20 --warmup 5 > gpurun_out/ab2_$v.json 2>gpurun_out/ab2_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab2_$v.json'))
print('$v rep$rep', round(d['value'],2), 'views/s', {k:round(x['avg_ms'],4) for k,x in d['roofline']['kernels'].items() if 'sample' in k})"
done; done
