# Third-session evidence on one B200: full GPU tests, default bench, reference arm, configs[4] end to end, then (only after
# those exited 0 without a profiler) the ncu launch list of the bench command and one --set full capture of the sweep kernels
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02s3_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02s3_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02s3_bench_1gpu.json 2> gpurun_out/r02s3_bench_1gpu.err || exit 1
timeout 600 python bench.py --impl reference > gpurun_out/r02s3_bench_ref.json 2> gpurun_out/r02s3_bench_ref.err
timeout 900 python bench.py --config full4k > gpurun_out/r02s3_full4k_1gpu.json 2> gpurun_out/r02s3_full4k_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_full4k_1gpu.json')); print('full4k', d['value'], d['seconds'], d['phase_seconds'])"
timeout 600 python bench.py --steps 1 --warmup 3 --views-per-gpu 1 --streams 1 --no-e2e --no-cpu-baseline --no-extras --no-plane-sweep > gpurun_out/r02s3_bench_1view.json 2>/dev/null || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02s3_launches_bench_1view.csv python bench.py --steps 1 --warmup 3 --views-per-gpu 1 --streams 1 --no-e2e --no-cpu-baseline --no-extras --no-plane-sweep > gpurun_out/ncu1.log 2>&1
timeout 300 python profiles/microbench/r02_sep_probe.py sweep > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ncc_sep_kernel|sample_kernel' --launch-skip 8 -c 4 -o gpurun_out/r02s3_sweep python profiles/microbench/r02_sep_probe.py sweep > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -12
