set -x
mkdir -p gpurun_out
for c in 64 256 512; do
python bench.py --steps 5 --warmup 3 --chunk $c --no-cpu-baseline --match-pairs 64 --allpairs-kf 0 --e2e-steps 2 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_c$c.json'))
print('chunk', $c, 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})
PY
done
