"""GPU parity of the stereo front-end (both eyes + Frame::ComputeStereoMatches, Frame.cc:831-1082)
through the C ABI against the CPU oracle: mvuRight / mvDepth bit-exact (the float arithmetic of the
sub-pixel fit is evaluated in the reference's order, without FMA)."""
import numpy as np
import pytest

from conftest import CONFIGS
from test_oracle_stereo import stereo_pair

pytestmark = pytest.mark.gpu


def _check_pair(oracle, gpu, left, right, nfeat, mbf, mb):
    eL = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7); eR = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    okl, odl = eL(left); okr, odr = eR(right)
    our, odp, n = oracle.stereo_matches(eL, eR, okl, odl, okr, odr, mbf, mb)
    kl, dl, kr, dr, ur, dp = gpu.extract_stereo(left, right, mbf, mb)
    assert len(kl) == len(okl) and len(kr) == len(okr)
    for f in ("x", "y", "octave", "response"):
        assert np.array_equal(kl[f], okl[f]) and np.array_equal(kr[f], okr[f])
    assert (dl == odl).all(1).mean() >= 0.999 and (dr == odr).all(1).mean() >= 0.999
    assert np.array_equal(ur.view(np.uint32), our.view(np.uint32)), "mvuRight differs"
    assert np.array_equal(dp.view(np.uint32), odp.view(np.uint32)), "mvDepth differs"
    return n, len(kl)


@pytest.mark.parametrize("name,disp", [("kitti", 17), ("euroc", 31), ("tum1", 4)])
def test_stereo_matches_oracle(oracle, name, disp):
    from orb_slam2_detailed_comments_b200 import ORBextractor
    w, h, nfeat = CONFIGS[name]
    gpu = ORBextractor(nfeat, 1.2, 8, 20, 7, max_batch=4)
    left, right = stereo_pair(w, h, 5, disp)
    n, nl = _check_pair(oracle, gpu, left, right, nfeat, 386.1448, 386.1448 / 718.856)  # KITTI00-02.yaml bf, b = bf/fx
    print(name, "stereo points", n, "of", nl)
    assert n > nl // 3


def test_stereo_edge_cases(oracle):
    from orb_slam2_detailed_comments_b200 import ORBextractor
    gpu = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=2)
    left, right = stereo_pair(640, 480, 9, 12)
    # unrelated right image: almost nothing survives; zero disparity: the 0.01 clamp path (:1051-1055)
    other = stereo_pair(640, 480, 77, 0)[0]
    _check_pair(oracle, gpu, left, other, 1000, 40.0, 0.1)
    _check_pair(oracle, gpu, left, left.copy(), 1000, 40.0, 0.1)
    # tiny disparity range: every candidate falls outside [uL - maxD, uL]
    _check_pair(oracle, gpu, left, right, 1000, 0.5, 0.1)
    # constant images: no keypoints at all
    flat = np.full((480, 640), 90, np.uint8)
    kl, dl, kr, dr, ur, dp = gpu.extract_stereo(flat, flat, 40.0, 0.1)
    assert len(kl) == 0 and len(kr) == 0 and len(ur) == 0


def test_stereo_batch_device(oracle):
    import torch
    from orb_slam2_detailed_comments_b200 import KP_DTYPE, ORBextractor
    w, h, nfeat = CONFIGS["kitti"]
    gpu = ORBextractor(nfeat, 1.2, 8, 20, 7, max_batch=4)   # 3 pairs through chunks of 2 pairs
    pairs = [stereo_pair(w, h, 20 + i, 10 + 7 * i) for i in range(3)]
    imgs = np.stack([im for p in pairs for im in p])
    cap = gpu.max_keypoints
    d_imgs = torch.from_numpy(imgs).cuda()
    d_kps = torch.zeros((6, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((6, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros(6, dtype=torch.int32, device="cuda")
    d_ur = torch.zeros((3, cap), dtype=torch.float32, device="cuda")
    d_dp = torch.zeros((3, cap), dtype=torch.float32, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    mbf, mb = 386.1448, 386.1448 / 718.856
    gpu.extract_stereo_batch_device(d_imgs, d_kps, d_desc, d_counts, d_ur, d_dp, mbf, mb, stream=ts.cuda_stream)
    gpu.synchronize(ts.cuda_stream)
    counts = d_counts.cpu().numpy(); ur = d_ur.cpu().numpy(); dp = d_dp.cpu().numpy()
    for p in range(3):
        eL = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7); eR = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
        okl, odl = eL(pairs[p][0]); okr, odr = eR(pairs[p][1])
        our, odp, n = oracle.stereo_matches(eL, eR, okl, odl, okr, odr, mbf, mb)
        assert counts[2 * p] == len(okl) and counts[2 * p + 1] == len(okr)
        assert np.array_equal(ur[p, :len(okl)].view(np.uint32), our.view(np.uint32))
        assert np.array_equal(dp[p, :len(okl)].view(np.uint32), odp.view(np.uint32))
        assert n > len(okl) // 3


@pytest.mark.parametrize("name,disp", [("kitti", 17), ("euroc", 31)])
def test_stereo_against_the_reference_itself(reference, name, disp):
    """Both eyes + ComputeStereoMatches on the GPU against the reference's own ORBextractor.cc + Frame.cc
    (oracle/_ref/liborbref.so, compiled unmodified): keypoints, descriptors, mvuRight and mvDepth bit-exact."""
    from orb_slam2_detailed_comments_b200 import ORBextractor
    w, h, nfeat = CONFIGS[name]
    gpu = ORBextractor(nfeat, 1.2, 8, 20, 7, max_batch=4)
    left, right = stereo_pair(w, h, 6, disp)
    rL = reference.ReferenceExtractor(nfeat, 1.2, 8, 20, 7); rR = reference.ReferenceExtractor(nfeat, 1.2, 8, 20, 7)
    rkl, rdl = rL(left); rkr, rdr = rR(right)
    mbf, mb = 386.1448, 386.1448 / 718.856
    rur, rdp, rn = reference.stereo_matches(rL, rR, mbf, mb)
    kl, dl, kr, dr, ur, dp = gpu.extract_stereo(left, right, mbf, mb)
    for f in ("x", "y", "size", "octave", "response", "class_id"):
        assert np.array_equal(kl[f], rkl[f]) and np.array_equal(kr[f], rkr[f]), f
    assert (dl == rdl).all(1).mean() >= 0.999 and (dr == rdr).all(1).mean() >= 0.999
    assert np.array_equal(ur.view(np.uint32), rur.view(np.uint32)), "mvuRight differs"
    assert np.array_equal(dp.view(np.uint32), rdp.view(np.uint32)), "mvDepth differs"
    print(name, "stereo points", rn, "of", len(kl))
    assert rn > len(kl) // 3
