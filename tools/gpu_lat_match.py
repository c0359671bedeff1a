import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from orb_slam2_detailed_comments_b200 import FrameView, ORBmatcher
from oracle import orb_oracle as O
from test_gpu_match_parity import _frames_from_extraction
ka, da, kb, db = _frames_from_extraction(O)
F1 = FrameView.from_keypoints(ka, da, 640, 480); F2 = FrameView.from_keypoints(kb, db, 640, 480)
m = ORBmatcher(0.9, True)
for mode in (0, 1):
    prev = F1.xy.copy()
    for _ in range(20): m.SearchForInitialization(F1, F2, prev.copy(), 100, mode=mode)
    t0 = time.perf_counter()
    for _ in range(300): m.SearchForInitialization(F1, F2, prev, 100, mode=mode)
    print("SearchForInitialization mode %d, %d x %d keypoints: %.3f ms per call" % (mode, F1.N, F2.N, (time.perf_counter() - t0) / 300 * 1e3))
