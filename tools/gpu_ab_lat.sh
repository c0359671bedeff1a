set -x
python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_multi_extractor.py tests/test_gpu_dropin_reference_frame.py -q -x -m gpu 2>&1 | tail -3
for v in 0 1 0 1; do echo "PDL=$v"; ORB_B200_PDL=$v python tools/gpu_lat1.py 2>&1 | grep -E "wall|host"; done
