set -x
mkdir -p gpurun_out
for w in 11 8 6 5 4 3; do
for t in 1 0; do
ORB_B200_FAST_TMA=$t ORB_B200_FAST_WARPS=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_w.json 2> gpurun_out/b_w.err
python -c "
import json; d=json.load(open('gpurun_out/b_w.json')); print('maxwarps $w tma $t', 'fps', d['value'], 'fast', d['stage_ms_per_step']['fast'])"
done
done
