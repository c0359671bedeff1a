# Collects the round's evidence on one B200: parity tests, smoke, bench lines (ours + reference arm), the ncu launch list of
# the bench command and ncu --set full captures of the extraction kernels and the matchers.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python -m pytest tests -m gpu -q -k "reference_itself or dropin or frame_constructor or initialisation_sequence" 2>&1 | tail -3 > gpurun_out/pytest_gpu_vs_reference.txt; cat gpurun_out/pytest_gpu_vs_reference.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python bench.py --workload euroc --no-cpu-baseline --allpairs-kf 0 --match-pairs 64 --sequence-frames 0 > gpurun_out/bench_euroc.json 2> gpurun_out/bench_euroc.err
python bench.py --workload tum1 --no-cpu-baseline --allpairs-kf 0 --match-pairs 64 --sequence-frames 0 > gpurun_out/bench_tum1.json 2> gpurun_out/bench_tum1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --pairs 256 --match-pairs 512 --allpairs-kf 64 --sequence-frames 0 --parity-frames 0 --no-cpu-baseline --no-latency > gpurun_out/ncu_launches.log 2>&1
ORB_B200_LANES=1 ncu --set full --clock-control none --import-source on -k regex:"k_level0_border2|k_resize_strip|k_fill_borders|k_fast_cells|k_quadtree|k_blur7|k_describe_ring|k_describe_tma" -s 39 -c 13 -o gpurun_out/prof_extract -f python bench.py --steps 1 --warmup 3 --pairs 128 --chunk 256 --match-pairs 64 --allpairs-kf 0 --sequence-frames 0 --parity-frames 0 --no-cpu-baseline --no-latency > gpurun_out/ncu_extract.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_match_pairs_bf|k_allpairs" -s 1 -c 2 -o gpurun_out/prof_match -f python bench.py --steps 1 --warmup 3 --pairs 64 --chunk 128 --match-pairs 1024 --allpairs-kf 128 --sequence-frames 0 --parity-frames 0 --no-cpu-baseline --no-latency > gpurun_out/ncu_match.log 2>&1
python tools/gpu_lat1.py > gpurun_out/latency.txt 2>&1; cat gpurun_out/latency.txt
ls -la gpurun_out | tail -15
python tools/gpu_dropin_latency.py > gpurun_out/dropin_latency.txt 2>&1; cat gpurun_out/dropin_latency.txt
