set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
for k in ('value','ms_per_step','e2e','stage_ms_per_step','path_roofline'):
    print(k, d[k])
PY
ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 12 -o gpurun_out/prof_all python bench.py --steps 1 --warmup 3 --pairs 64 --match-pairs 64 --no-cpu-baseline > gpurun_out/ncu_all.log 2>&1
ls -la gpurun_out
