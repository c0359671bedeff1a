set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 12 -o gpurun_out/prof_all3 python bench.py --steps 1 --warmup 3 --pairs 64 --chunk 128 --match-pairs 64 --allpairs-kf 0 --no-cpu-baseline --no-latency > gpurun_out/ncu_all2.log 2>&1
ls -la gpurun_out
