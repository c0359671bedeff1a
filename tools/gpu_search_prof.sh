set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 > gpurun_out/bench_tr.json 2> gpurun_out/bench_tr.err; tail -3 gpurun_out/bench_tr.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tr.json')); print(d['value'], d['tracking'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_search|k_project|k_grid_assign|k_bow' -c 8 -o gpurun_out/prof_search python bench.py --steps 1 --warmup 3 --pairs 64 --match-pairs 64 --no-cpu-baseline --no-latency --allpairs-kf 0 > gpurun_out/ncu_search.log 2>&1
ls -la gpurun_out | tail -5
