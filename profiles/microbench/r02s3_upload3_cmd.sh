# same-box A/B at the full 200 views of configs[4]: driver's pageable upload path vs the opt-in page-locked ring
for m in direct ring direct ring; do
B200D_UPLOAD=$m timeout 900 python bench.py --config full4k > gpurun_out/r02s3_full4k200_$m.json 2> gpurun_out/r02s3_full4k200_$m.err
python -c "
import json; d=json.load(open('gpurun_out/r02s3_full4k200_$m.json')); print('$m', round(d['value'],2), round(d['seconds']['total'],3), round(d['phase_seconds']['depth_maps_incl_upload_prepare'],3), d['host_side_seconds_inside_depth_maps'])"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv; nproc; free -g | head -2
