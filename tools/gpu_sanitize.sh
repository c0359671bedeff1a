set -x
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py tests/test_gpu_match_parity.py -m gpu -q -x -k "tum1 or euroc or stereo_edge or adversarial or dedup or reference_mode or allpairs or batch_device or small_and_odd" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck exit: $?" >> gpurun_out/sanitize_memcheck.txt
tail -25 gpurun_out/sanitize_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py -m gpu -q -x -k "tum1 or stereo_edge" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck exit: $?" >> gpurun_out/sanitize_racecheck.txt
tail -25 gpurun_out/sanitize_racecheck.txt
