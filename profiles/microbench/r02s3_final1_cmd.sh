# Third-session final evidence on one B200: full GPU tests, smoke, default bench, reference arm, configs[4] end to end
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02s3_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02s3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02s3_bench_1gpu.json 2> gpurun_out/r02s3_bench_1gpu.err || exit 1
timeout 600 python bench.py --impl reference > gpurun_out/r02s3_bench_ref.json 2> gpurun_out/r02s3_bench_ref.err
timeout 900 python bench.py --config full4k > gpurun_out/r02s3_full4k_1gpu.json 2> gpurun_out/r02s3_full4k_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['plane_sweep']['value'], d['plane_sweep']['e2e']['value'], d['e2e_reconstruct']['value'])
d=json.load(open('gpurun_out/r02s3_full4k_1gpu.json')); print('full4k', d['value'], d['seconds'], d['phase_seconds'], d['host_side_seconds_inside_depth_maps'])"
