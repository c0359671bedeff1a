set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_extract_parity.py -q -x 2>&1 | tail -3
for c in 256 128; do
timeout 300 python bench.py --steps 5 --warmup 3 --chunk $c --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_$c.json 2> gpurun_out/b_$c.err
python -c "
import json; d=json.load(open('gpurun_out/b_$c.json')); print($c, 'fps', d['value'], 'e2e', d['e2e']['value'], d['stage_ms_per_step'])"
done
