# (record of a rejected experiment: B200D_UPLOAD=ring selected a page-locked staging ring in DenseEngine.to_device; see r02s3_upload_ab.py)
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02s3_bench_1gpu_b.json 2> gpurun_out/r02s3_bench_1gpu_b.err || exit 1
python -c "
import json; d=json.load(open('gpurun_out/r02s3_bench_1gpu_b.json')); print(d['value'], d['e2e']['value'], d['plane_sweep']['value'], d['plane_sweep']['e2e']['value'], d['e2e_reconstruct']['value'], d['e2e_reconstruct']['phases_rank0_s'])"
for m in ring direct ring direct; do
B200D_UPLOAD=$m timeout 900 python bench.py --config full4k --full4k-views 100 > gpurun_out/r02s3_full4k100_$m.json 2> gpurun_out/r02s3_full4k100_$m.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_full4k100_$m.json')); print('$m', round(d['value'],2), d['seconds']['total'], d['phase_seconds']['depth_maps_incl_upload_prepare'], d['host_side_seconds_inside_depth_maps'])"
done
