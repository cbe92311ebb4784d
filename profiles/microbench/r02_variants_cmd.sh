# Round-2 variant scans (run on the GPU box; B200DENSE_LIB selects a variant .so built by build_variants.py):
#  1. separable NCC kernel, fast PatchMatch modes: 3 vs 2 resident CTAs/SM -> 209 vs 253 views/s (2 shipped); 1 CTA/SM: 181
#     python profiles/microbench/build_variants.py pm1:B200D_SEP_MIN_BLOCKS_PM=1 ; --arith fast --streams 4/6
#  2. tile height 16 (B200D_TH=16,B200D_MIN_BLOCKS=6,B200D_SEP_MIN_BLOCKS_SWEEP=6): sweep NCC 75.4 -> 72.6 us, parity PatchMatch 181 -> 155
#  3. chain kernel limited to 2 CTAs/SM by 38 KB of shared-memory padding (80 registers kept): isolated launch 158 us (= 3 CTAs),
#     4-stream step 152 views/s (vs 180); 20 KB padding (3 CTAs, smaller L1): 179-181
#  4. views in flight: --streams 3/6/8 -> 180.7 / 181.3 / 181.8 ; --sweep-streams 2/3/4 -> 20.1 / 19.1 / 20.0
for st in 4 6; do
  python bench.py --arith fast --no-cpu-baseline --no-extras --no-e2e --no-plane-sweep --steps 3 --streams $st > gpurun_out/varf.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/varf.json')); print('fast streams $st', round(d['value'],1), {k:round(x['avg_ms'],4) for k,x in d['roofline']['kernels'].items() if 'REFINE' in k})"
done
