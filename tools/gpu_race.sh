export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py tests/test_gpu_search_parity.py tests/test_gpu_input_parity.py -m gpu -q -x -k "tum1 or stereo_edge or last_frame or bow or triangulation or distinctive or (cvt_gray and 640)" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck exit: $?" >> gpurun_out/sanitize_racecheck.txt
tail -12 gpurun_out/sanitize_racecheck.txt
