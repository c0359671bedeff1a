set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --pairs 256 --match-pairs 512 --allpairs-kf 64 --no-cpu-baseline --no-latency > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
