"""Per-stage device time of ONE frame per call (the drop-in path's latency), KITTI / TUM1 shape."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from orb_slam2_detailed_comments_b200 import ORBextractor
from orb_slam2_detailed_comments_b200.synth import synth_frame

for (w, h, nf) in ((1241, 376, 2000), (640, 480, 1000)):
    ext = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=1)
    img = synth_frame(w, h, 5)
    for _ in range(20):
        ext(img)
    ext.set_profiling(True)
    ext.stage_times()
    n = 50
    for _ in range(n):
        ext(img)
    st = ext.stage_times()
    ext.set_profiling(False)
    t0 = time.perf_counter()
    for _ in range(200):
        ext(img)
    wall = (time.perf_counter() - t0) / 200 * 1e3
    print(w, h, "wall ms/call %.3f" % wall, {k: round(v[0] / n * 1e3, 1) for k, v in st.items()}, "us per call (device, per stage)")
    acc = {}
    for _ in range(100):
        ext(img)
        for k, v in ext.last_call_breakdown().items():
            acc[k] = acc.get(k, 0.0) + v / 100
    print("   host breakdown (us):", {k: round(v, 1) for k, v in acc.items()})
