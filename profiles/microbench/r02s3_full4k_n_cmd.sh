# BASELINE configs[4] end to end at N GPUs (N = first argument), final third-session code (block sharding, staged uploads, exact PLY formatter)
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --config full4k > gpurun_out/r02s3_full4k_${N}gpu.json 2> gpurun_out/r02s3_full4k_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_full4k_${N}gpu.json')); print(d['value'], d['seconds'], d['host_side_seconds_inside_depth_maps'], d['phase_seconds'], d['points_out'], d['ply_bytes'])"
