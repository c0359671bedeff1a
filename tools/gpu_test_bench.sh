set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
for k in ('value','ms_per_step','e2e','stage_ms_per_step','roofline','path_roofline','clocks'):
    print(k, d[k])
m=d['matching']; print({k:m[k] for k in ('value','ms_per_step')}, m['roofline']['frac'])
PY
