set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_stereo -s 1 -c 1 -o gpurun_out/prof_stereo python bench.py --steps 1 --warmup 3 --pairs 296 --chunk 592 --match-pairs 64 --allpairs-kf 0 --no-cpu-baseline > gpurun_out/ncu_stereo.log 2>&1
ls -la gpurun_out | tail -3
