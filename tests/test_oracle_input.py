"""Pins the oracle's input stage against the real OpenCV available here (cv2): cvtColor to gray for the four
channel orders Tracking.cc uses, remap INTER_LINEAR with float maps (stereo_euroc.cc rectification), and
ComputeDistinctiveDescriptors against a direct Python restatement of MapPoint.cc:365-448."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("code,channels,blue_first", [(cv2.COLOR_RGB2GRAY, 3, False), (cv2.COLOR_BGR2GRAY, 3, True),
                                                      (cv2.COLOR_RGBA2GRAY, 4, False), (cv2.COLOR_BGRA2GRAY, 4, True)])
def test_cvt_gray_matches_cv2(oracle, code, channels, blue_first):
    rng = np.random.RandomState(channels + blue_first)
    img = rng.randint(0, 256, (123, 217, channels)).astype(np.uint8)
    assert np.array_equal(oracle.cvt_gray(img, blue_first), cv2.cvtColor(img, code))
    # every (r, g) pair with a strided b: the rounding of all coefficient sums
    full = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(0, 256, 5), indexing="ij"), -1).reshape(256, -1, 3).astype(np.uint8)
    if channels == 3:
        assert np.array_equal(oracle.cvt_gray(full, blue_first), cv2.cvtColor(full, code))


def rectify_maps(w, h, seed):
    """EuRoC-like rectification maps from cv2.initUndistortRectifyMap with a distorted camera and a small rotation."""
    rng = np.random.RandomState(seed)
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]]) * (w / 752.0); K[2, 2] = 1
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    R, _ = cv2.Rodrigues(rng.randn(3) * 0.01)
    P = K.copy(); P[0, 2] += 3.0
    return cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32F)


@pytest.mark.parametrize("w,h", [(752, 480), (333, 211)])
def test_remap_matches_cv2(oracle, w, h):
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    img = synth_frame(w, h, 7)
    mx, my = rectify_maps(w, h, 3)
    assert np.array_equal(oracle.remap_linear(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    # maps that leave the image on every side, exact integer coordinates, exact halves
    rng = np.random.RandomState(1)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    mx2 = (xx * 1.1 - 20 + np.round(rng.rand(h, w) * 64) / 32).astype(np.float32)
    my2 = (yy * 1.1 - 15 + np.round(rng.rand(h, w) * 4) / 2).astype(np.float32)
    assert np.array_equal(oracle.remap_linear(img, mx2, my2), cv2.remap(img, mx2, my2, cv2.INTER_LINEAR))
    noise = rng.randint(0, 256, (h, w)).astype(np.uint8)
    mx3 = (rng.rand(h, w) * (w + 8) - 4).astype(np.float32); my3 = (rng.rand(h, w) * (h + 8) - 4).astype(np.float32)
    assert np.array_equal(oracle.remap_linear(noise, mx3, my3), cv2.remap(noise, mx3, my3, cv2.INTER_LINEAR))


def test_distinctive_descriptors(oracle):
    rng = np.random.RandomState(9)
    counts = [1, 2, 3, 4, 7, 16, 31, 32, 33, 64, 100, 0, 5]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    desc = rng.randint(0, 256, (offsets[-1], 32)).astype(np.uint8)
    # clustered observations (realistic) and duplicates (ties)
    for p, c in enumerate(counts):
        if c > 2:
            base = desc[offsets[p]].copy()
            for i in range(c):
                d = base.copy()
                bits = rng.randint(0, 256, rng.randint(0, 40))
                np.bitwise_xor.at(d, bits >> 3, (1 << (bits & 7)).astype(np.uint8))
                desc[offsets[p] + i] = d
            desc[offsets[p] + c - 1] = desc[offsets[p]]
    got = oracle.distinctive_descriptors(desc, offsets)
    pop = np.array([bin(i).count("1") for i in range(256)])
    for p, c in enumerate(counts):
        if c == 0:
            assert got[p] == -1
            continue
        d = desc[offsets[p]:offsets[p + 1]]
        dist = pop[d[:, None, :] ^ d[None, :, :]].sum(-1)
        best, best_idx = 1 << 30, 0
        for i in range(c):
            med = sorted(dist[i].tolist())[int(0.5 * (c - 1))]
            if med < best:
                best, best_idx = med, i
        assert got[p] == best_idx
