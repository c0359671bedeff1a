import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_detailed_comments_b200 import ORBextractor
from orb_slam2_detailed_comments_b200.synth import synth_frame
img = synth_frame(640, 480, 3)
ext = ORBextractor(1000, 1.2, 8, 20, 7, device=0, max_batch=4)
k, d = ext(img)
print("keypoints", len(k), d[:2].tolist() if len(k) else None)
