import sys, time, numpy as np
sys.path.insert(0, ".")
from orb_slam2_detailed_comments_b200 import ORBextractor
from orb_slam2_detailed_comments_b200.synth import synth_frame
ext = ORBextractor(2000, 1.2, 8, 20, 7, max_batch=1)
img = synth_frame(1241, 376, 5)
for _ in range(20): ext(img, want_pyramid=True)
t0 = time.perf_counter()
for _ in range(200): ext(img, want_pyramid=True)
print("with pyramid wall ms/call %.3f" % ((time.perf_counter() - t0) / 200 * 1e3))
