mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py tests/test_abi.py -q -x 2>&1 | tail -3
python tools/gpu_lat1.py
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from orb_slam2_detailed_comments_b200 import ORBextractor
from test_oracle_stereo import stereo_pair
ext = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=2)
l, r = stereo_pair(1241, 376, 5, 17)
for _ in range(20): ext.extract_stereo(l, r, 386.1448, 0.5371)
t0 = time.perf_counter()
for _ in range(200): ext.extract_stereo(l, r, 386.1448, 0.5371)
print("stereo pair wall ms/call %.3f" % ((time.perf_counter() - t0) / 200 * 1e3))
PY
