# second-session check: GPU parity suite (incl. the test against the reference's own ORBextractor.cc),
# smoke, reference arm, default bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python -m pytest tests/test_gpu_extract_parity.py -m gpu -q -s -k reference_itself 2>&1 | grep -v "^$" | tail -14 > gpurun_out/pytest_gpu_ref.txt; cat gpurun_out/pytest_gpu_ref.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
for k in ('value','ms_per_step','e2e','stage_ms_per_step','roofline','path_roofline','clocks','cpu_baseline'):
    print(k, d[k])
PY
