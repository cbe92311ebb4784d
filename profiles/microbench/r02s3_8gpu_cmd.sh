# 8-GPU evidence of the third session: default bench line (block sharding, packed sweep kernels) + BASELINE configs[4] end to end
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02s3_bench_8gpu.json 2> gpurun_out/r02s3_bench_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_bench_8gpu.json')); print(d['value'], d['e2e']['value'], d['gather_check']['ok'], d['gather_check']['points'], d['plane_sweep']['value'], d['plane_sweep']['e2e']['value'], d['e2e_reconstruct']['value'], d['exchange']['all_gather_gbs_per_rank'], d['fast_mode']['value'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config full4k > gpurun_out/r02s3_full4k_8gpu.json 2> gpurun_out/r02s3_full4k_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_full4k_8gpu.json')); print(d['value'], d['seconds'], d['host_side_seconds_inside_depth_maps'], d['phase_seconds'], d['points_out'], d['ply_bytes'])"
