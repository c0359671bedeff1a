# quick A/B: selected parity tests + a short device-resident bench (no CPU baseline, no matching section)
# usage: gpurun -- bash tools/gpu_q.sh [pytest targets...]
set -x
mkdir -p gpurun_out
T=${@:-tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py}
timeout 1200 python -m pytest $T -q -x -m gpu 2>&1 | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_q.json 2> gpurun_out/b_q.err; tail -2 gpurun_out/b_q.err
python -c "
import json; d=json.load(open('gpurun_out/b_q.json')); print('fps', d['value'], 'e2e', d['e2e']['value'], d['stage_ms_per_step'])"
