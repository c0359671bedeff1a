# quick A/B: extractor parity tests + a short device-resident bench (no CPU baseline, no matching section)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_q.json 2> gpurun_out/b_q.err; tail -2 gpurun_out/b_q.err
python -c "
import json; d=json.load(open('gpurun_out/b_q.json')); print('fps', d['value'], 'e2e', d['e2e']['value'], d['stage_ms_per_step'])"
