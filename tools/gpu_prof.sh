set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --pairs 64 --match-pairs 256 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fast_cells -s 3 -c 1 -o gpurun_out/prof_fast python bench.py --steps 1 --warmup 3 --pairs 64 --match-pairs 64 --no-cpu-baseline > gpurun_out/ncu_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_match_pairs -s 1 -c 1 -o gpurun_out/prof_match python bench.py --steps 1 --warmup 3 --pairs 64 --match-pairs 256 --no-cpu-baseline > gpurun_out/ncu_match.log 2>&1
ls -la gpurun_out
