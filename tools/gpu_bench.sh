set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -20 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
