set -x
mkdir -p gpurun_out
for cfg in "2 1" "1 1" "2 2" "4 2" "1 2" "2 4"; do
set -- $cfg
ORB_B200_FAST_TAIL_RUN=$1 ORB_B200_FAST_TAIL_MUL=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_t.json 2> gpurun_out/b_t.err
python -c "
import json; d=json.load(open('gpurun_out/b_t.json')); print('run $1 mul $2', 'fps', d['value'], 'fast', d['stage_ms_per_step']['fast'])"
done
