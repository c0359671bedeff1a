# A/B of ORB_B200_BLUR_FORK (blur on a side stream beside quadtree / FAST+quadtree)
mkdir -p gpurun_out
for v in 0 2 1 0 2 1; do
  ORB_B200_BLUR_FORK=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_fork$v.json 2> gpurun_out/b_fork$v.err
  python -c "
import json; d=json.load(open('gpurun_out/b_fork$v.json')); print('fork $v fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'stereo', round(d['stereo']['value']))"
done
ORB_B200_BLUR_FORK=2 timeout 600 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py tests/test_gpu_full_size.py -q -x 2>&1 | tail -2
