"""CPU stress of the pinning: the oracle against the reference's own ORBextractor.cc (oracle/_ref, canonical heap) on many
random frames, shapes and parameters. usage: python tools/ref_stress.py [frames]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import orb_oracle as O, orb_ref as R
from orb_slam2_detailed_comments_b200.synth import synth_frame

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 2026)
bad = 0; kp = 0; t0 = time.time()
for i in range(n):
    w = int(rng.randint(160, 1300)); h = int(rng.randint(120, 520))
    # landscape only: for portrait images the reference computes nIni = round(width / height) = 0 roots at the small
    # levels and then indexes an empty vector (ORBextractor.cc:695-739) - it crashes (seen here with 168 x 275, scale
    # 1.1); the product returns ORB_ERR_UNSUPPORTED for such sizes
    if w < h:
        w, h = h, w
    w = min(w, 3 * h)
    nf = int(rng.choice([200, 500, 1000, 1200, 2000, 3000]))
    nl = int(rng.choice([8, 8, 8, 5, 10])); sf = float(rng.choice([1.2, 1.2, 1.1, 1.3, 1.5]))
    ini = int(rng.choice([20, 20, 12, 30])); mn = int(rng.choice([7, 7, 5, 10]))
    # the top level must still hold one 30-px cell
    while nl > 1 and min(w, h) / sf ** (nl - 1) < 70:
        nl -= 1
    kind = rng.randint(4)
    if kind == 0:
        img = rng.randint(0, 256, (h, w)).astype(np.uint8)
    else:
        img = synth_frame(w, h, 1000 + i, n_rect=int(rng.randint(50, 800)), noise_sigma=float(rng.rand() * 6))
    if '-v' in sys.argv: print(i, w, h, nf, nl, sf, ini, mn, kind, flush=True)
    o = O.OracleExtractor(nf, sf, nl, ini, mn); r = R.ReferenceExtractor(nf, sf, nl, ini, mn)
    ko, do = o(img); kr, dr = r(img)
    ok = len(ko) == len(kr) and ko.tobytes() == kr.tobytes() and np.array_equal(do, dr) and \
        all(np.array_equal(o.level(l), r.level(l)) for l in range(nl))
    kp += len(ko)
    if not ok:
        bad += 1
        print("MISMATCH", i, w, h, nf, nl, sf, ini, mn, kind, len(ko), len(kr))
print("%d frames, %d keypoints, %d mismatches, %.1f s" % (n, kp, bad, time.time() - t0))
sys.exit(1 if bad else 0)
