# 2-GPU validation after the switch to block sharding: multi-GPU parity tests + N=2 bench line (gather_check) + full4k at N=2
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02s3_pytest_multi_2gpu.log 2>&1; tail -4 gpurun_out/r02s3_pytest_multi_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02s3_bench_2gpu.json 2> gpurun_out/r02s3_bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_bench_2gpu.json')); print(d['value'], d['e2e']['value'], d['gather_check'], d['plane_sweep']['value'], d['e2e_reconstruct']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config full4k > gpurun_out/r02s3_full4k_2gpu.json 2> gpurun_out/r02s3_full4k_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_full4k_2gpu.json')); print(d['value'], d['seconds'], d['host_side_seconds_inside_depth_maps'], d['points_out'], d['ply_bytes'])"
