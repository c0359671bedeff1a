"""Pins the oracle's ComputeStereoMatches against the independent Python/cv2 restatement."""
import numpy as np
import pytest

pytest.importorskip("cv2")

from stereo_reference import compute_stereo_matches  # noqa: E402
from orb_slam2_detailed_comments_b200.synth import synth_frame  # noqa: E402


def stereo_pair(w, h, seed, disparity=11):
    big = synth_frame(w + 64, h, seed, noise_sigma=0).astype(np.float32)
    rng = np.random.RandomState(seed)
    left = np.clip(np.rint(big[:, 32:32 + w] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    right = np.clip(np.rint(big[:, 32 + disparity:32 + disparity + w] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    return left, right


@pytest.mark.parametrize("w,h,nfeat,disp", [(640, 480, 800, 9), (752, 480, 1200, 23)])
def test_stereo_oracle_vs_python(oracle, w, h, nfeat, disp):
    left, right = stereo_pair(w, h, 3, disp)
    eL = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7); eR = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    kl, dl = eL(left); kr, dr = eR(right)
    mbf, mb = 40.0, 0.1
    ur, dp, n = oracle.stereo_matches(eL, eR, kl, dl, kr, dr, mbf, mb)
    ref_ur, ref_dp = compute_stereo_matches([eL.level(l) for l in range(8)], [eR.level(l) for l in range(8)],
                                            eL.scale, eL.inv_scale, kl, dl, kr, dr, mbf, mb)
    assert np.array_equal(ur.view(np.uint32), ref_ur.view(np.uint32))
    assert np.array_equal(dp.view(np.uint32), ref_dp.view(np.uint32))
    assert n == int((ur >= 0).sum()) > len(kl) // 3
    good = ur >= 0
    assert abs(np.median((kl["x"] - ur)[good]) - disp) < 0.5
