# full4k (64 views) at 2 GPUs: peer-store gather vs NCCL point-to-point gather -- where do fusion and filter time go?
for mode in peer nccl; do
  if [ $mode = nccl ]; then export B200D_NO_PEER_GATHER=1; else unset B200D_NO_PEER_GATHER; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config full4k --full4k-views 64 > gpurun_out/r02s3_full4k64_$mode.json 2> gpurun_out/r02s3_full4k64_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/r02s3_full4k64_$mode.json')); print('$mode', round(d['value'],1), d['seconds'], d['phase_seconds'], d['raw_points'])"
done
