set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_fast_cells|k_describe" -s 6 -c 2 -o gpurun_out/prof_two python bench.py --steps 1 --warmup 3 --pairs 64 --match-pairs 64 --no-cpu-baseline > gpurun_out/ncu_two.log 2>&1
ls -la gpurun_out
