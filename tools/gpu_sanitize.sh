set -x
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py tests/test_gpu_match_parity.py tests/test_gpu_frame_parity.py tests/test_gpu_search_parity.py tests/test_gpu_input_parity.py tests/test_gpu_multi_extractor.py tests/test_gpu_dropin_reference_frame.py -m gpu -q -x -k "tum1 or euroc or stereo_edge or adversarial or dedup or reference_mode or allpairs or batch_device or small_and_odd or undistort or last_frame or local_map or bow or triangulation or cvt_gray or remap or distinctive or budgets or batch_and_single or two_threads or monocular or stereo_frame or latency_form or through_the_reference_class" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck exit: $?" >> gpurun_out/sanitize_memcheck.txt
tail -12 gpurun_out/sanitize_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py tests/test_gpu_search_parity.py tests/test_gpu_input_parity.py tests/test_gpu_multi_extractor.py tests/test_gpu_match_parity.py -m gpu -q -x -k "tum1 or stereo_edge or budgets or batch_and_single or last_frame or bow or triangulation or distinctive or (cvt_gray and 640) or (latency_form and (tiny or clustered_wide or extraction_narrow))" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck exit: $?" >> gpurun_out/sanitize_racecheck.txt
tail -12 gpurun_out/sanitize_racecheck.txt
