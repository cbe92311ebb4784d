set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02s3_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02s3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r02s3_bench_1gpu.json 2> gpurun_out/r02s3_bench_1gpu.err ) 2>&1 | tail -3
( time timeout 600 python bench.py --impl reference > gpurun_out/r02s3_bench_ref.json 2> gpurun_out/r02s3_bench_ref.err ) 2>&1 | tail -3
for ss in 2 3; do timeout 300 python bench.py --no-cpu-baseline --no-extras --no-e2e --steps 2 --sweep-streams $ss > gpurun_out/r02s3_sw$ss.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02s3_sw$ss.json')); print('sweep-streams $ss', d['plane_sweep']['value'], d['plane_sweep']['rep_ms'])"; done
