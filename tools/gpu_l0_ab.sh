mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py tests/test_gpu_full_size.py -q -x 2>&1 | tail -3
for v in 1 0 1 0; do
  ORB_B200_L0_FORK=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_l0$v.json 2> gpurun_out/b_l0$v.err
  python -c "
import json; d=json.load(open('gpurun_out/b_l0$v.json')); print('l0fork $v fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'stereo', round(d['stereo']['value']))"
done
export PATH=/usr/local/cuda/bin:$PATH
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_extract_parity.py -m gpu -q -x -k "tum1 or adversarial or small_and_odd or reference_itself" 2>&1 | tail -6
