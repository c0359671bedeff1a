# ncu --set full capture of kernels matching a regex inside a small bench run
# usage: gpurun -- bash tools/gpu_prof.sh <kernel regex> <skip> <count> <out name> [extra bench args]
set -x
mkdir -p gpurun_out
K=$1; S=$2; C=$3; O=$4; shift 4
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -o gpurun_out/$O -f python bench.py --steps 1 --warmup 3 --pairs 128 --chunk 256 --match-pairs 64 --allpairs-kf 0 --no-cpu-baseline --no-latency "$@" > gpurun_out/ncu_$O.log 2>&1
tail -3 gpurun_out/ncu_$O.log
ls -la gpurun_out/$O.ncu-rep
