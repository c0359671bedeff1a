"""CPU stress: oracle vs the reference's own Frame::ComputeStereoMatches on random stereo pairs."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import orb_oracle as O, orb_ref as R
from test_oracle_stereo import stereo_pair
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(7)
bad = 0; pts = 0; t0 = time.time()
for i in range(n):
    w = int(rng.randint(320, 1300)); h = int(rng.randint(240, 500)); w = max(w, h)
    nf = int(rng.choice([500, 1000, 2000])); disp = int(rng.randint(0, 31))
    left, right = stereo_pair(w, h, 300 + i, disp)
    if rng.rand() < 0.1:
        right = stereo_pair(w, h, 900 + i, 0)[0]          # unrelated right image
    eL = O.OracleExtractor(nf); eR = O.OracleExtractor(nf); rL = R.ReferenceExtractor(nf); rR = R.ReferenceExtractor(nf)
    kl, dl = eL(left); kr, dr = eR(right); rL(left); rR(right)
    mbf = float(rng.choice([386.1448, 40.0, 0.5])); mb = float(rng.choice([0.5371, 0.1]))
    ur, dp, k = O.stereo_matches(eL, eR, kl, dl, kr, dr, mbf, mb)
    rur, rdp, rk = R.stereo_matches(rL, rR, mbf, mb)
    ok = np.array_equal(ur.view(np.uint32), rur.view(np.uint32)) and np.array_equal(dp.view(np.uint32), rdp.view(np.uint32)) and k == rk
    pts += k; bad += not ok
    if not ok: print("MISMATCH", i, w, h, nf, disp, mbf, mb)
print("%d pairs, %d stereo points, %d mismatching pairs, %.1f s" % (n, pts, bad, time.time() - t0))
sys.exit(1 if bad else 0)
