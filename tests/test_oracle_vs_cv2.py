"""Pins the CPU oracle: every OpenCV primitive it restates is checked bit-exactly against the
real OpenCV (python cv2 4.13.0), and the whole extractor against the cv2-driven restatement
of the reference control flow (tests/cv2_reference.py) and the committed golden vectors."""
import glob
import os
import zlib

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from conftest import CONFIGS  # noqa: E402
from cv2_reference import Cv2Reference, distribute_quadtree, load_pattern  # noqa: E402
from orb_slam2_detailed_comments_b200.synth import adversarial_frames, synth_frame  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cv2_version_is_the_pinned_one():
    # blur/resize arithmetic is release dependent: the oracle pins 4.13.0 behaviour
    assert cv2.__version__.startswith("4.")


@pytest.mark.parametrize("name", list(CONFIGS))
def test_resize_chain_and_border(oracle, name):
    w, h, nfeat = CONFIGS[name]
    orc = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    img = synth_frame(w, h, 3)
    cur = img
    for l in range(1, 8):
        dw = int(np.rint(np.float32(w) * orc.inv_scale[l])); dh = int(np.rint(np.float32(h) * orc.inv_scale[l]))
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        got = oracle.resize_linear(cur, dw, dh)
        assert np.array_equal(got, ref), (name, l)
        assert np.array_equal(oracle.border101(got), cv2.copyMakeBorder(ref, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
        cur = ref


def test_resize_random_sizes(oracle):
    rng = np.random.RandomState(0)
    for _ in range(25):
        sw, sh = rng.randint(40, 400), rng.randint(40, 300)
        f = 1.05 + rng.rand() * 0.6
        dw, dh = max(8, int(round(sw / f))), max(8, int(round(sh / f)))
        src = rng.randint(0, 256, (sh, sw)).astype(np.uint8)
        for ipp in (True, False):
            cv2.ipp.setUseIPP(ipp)
            assert np.array_equal(oracle.resize_linear(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))
    cv2.ipp.setUseIPP(True)


@pytest.mark.parametrize("threshold", [20, 7])
def test_fast_on_subimage_views(oracle, threshold):
    rng = np.random.RandomState(threshold)
    det = cv2.FastFeatureDetector_create(threshold, True)
    det_nonms = cv2.FastFeatureDetector_create(threshold, False)
    imgs = [synth_frame(320, 240, 5), rng.randint(0, 256, (240, 320)).astype(np.uint8),
            adversarial_frames(320, 240)["checkerboard"]]
    n_checked = 0
    for img in imgs:
        for _ in range(120):
            sw, sh = rng.randint(7, 60), rng.randint(7, 60)
            x0, y0 = rng.randint(0, 320 - sw), rng.randint(0, 240 - sh)
            sub = img[y0:y0 + sh, x0:x0 + sw]
            for d, nms in ((det, True), (det_nonms, False)):
                kps = d.detect(sub)
                xs, ys, sc = oracle.fast(sub, threshold, nms)
                if nms:
                    ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps]
                    assert ref == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
                else:  # cv::FAST leaves response 0 when NMS is off: compare positions only
                    ref = [(int(k.pt[0]), int(k.pt[1])) for k in kps]
                    assert ref == list(zip(xs.tolist(), ys.tolist()))
                n_checked += len(ref)
    assert n_checked > 500


def test_fast_score_is_threshold_independent(oracle):
    img = synth_frame(200, 150, 9)
    S = oracle.fast_score_map(img)
    for t in (7, 20, 35):
        corners = {(int(k.pt[0]), int(k.pt[1])) for k in cv2.FastFeatureDetector_create(t, False).detect(img)}
        ys, xs = np.nonzero(S >= t)
        assert {(int(x), int(y)) for x, y in zip(xs, ys)} == corners          # corner at t  <=>  S >= t
        for k in cv2.FastFeatureDetector_create(t, True).detect(img):         # response == S for every t
            assert int(k.response) == int(S[int(k.pt[1]), int(k.pt[0])])


def test_gaussian_blur(oracle):
    rng = np.random.RandomState(1)
    for (w, h) in ((64, 48), (333, 217), (9, 9), (1241, 376)):
        for img in (rng.randint(0, 256, (h, w)).astype(np.uint8), synth_frame(w, h, 4)):
            ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            assert np.array_equal(oracle.gauss7(img), ref)


def test_fast_atan2(oracle):
    rng = np.random.RandomState(2)
    y = rng.randint(-200000, 200000, 30000).astype(np.float32)
    x = rng.randint(-200000, 200000, 30000).astype(np.float32)
    y[:100] = 0; x[50:150] = 0
    got = oracle.fast_atan2(y, x)
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert got[60] == 0.0  # fastAtan2(0, 0) = 0


def test_quadtree_against_python_restatement(oracle):
    rng = np.random.RandomState(5)
    for trial in range(30):
        W = int(rng.randint(100, 1300)); H = int(rng.randint(90, 500))
        if W < H // 2 + 1:
            continue
        n = int(rng.randint(1, 3000))
        pts = set()
        while len(pts) < n:
            if rng.rand() < 0.5:   # clustered
                cx, cy = rng.randint(3, W - 3), rng.randint(3, H - 3)
                x = int(np.clip(cx + rng.normal(0, 8), 3, W - 4)); y = int(np.clip(cy + rng.normal(0, 8), 3, H - 4))
            else:
                x, y = int(rng.randint(3, W - 3)), int(rng.randint(3, H - 3))
            pts.add((x, y))
        pts = sorted(pts, key=lambda p: (p[1], p[0]))
        score = rng.randint(7, 60, len(pts))
        N = int(rng.randint(5, 600))
        cand = [(float(x), float(y), float(s)) for (x, y), s in zip(pts, score)]
        ref = distribute_quadtree(cand, 16, 16 + W, 16, 16 + H, N)
        got, _ = oracle.quadtree([c[0] for c in cand], [c[1] for c in cand], score, 16, 16 + W, 16, 16 + H, N)
        assert got.tolist() == ref, trial


@pytest.mark.parametrize("name", ["tum1", "kitti"])
def test_extractor_end_to_end_vs_cv2_reference(oracle, name):
    w, h, nfeat = CONFIGS[name]
    img = synth_frame(w, h, 7)
    ref = Cv2Reference(nfeat, 1.2, 8, 20, 7, load_pattern())
    rk, rd, levels, dbg = ref(img)
    orc = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    kps, desc = orc(img)
    assert orc.per_level.tolist() == ref.per_level and orc.umax.tolist() == ref.umax
    assert np.array_equal(orc.scale, np.array(ref.scale, np.float32))
    for l in range(8):
        assert np.array_equal(orc.level(l), levels[l])
        xs, ys, sc = orc.candidates(l)
        assert [(int(c[0]) + 16, int(c[1]) + 16, int(c[2])) for c in dbg[l]["cand"]] == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
        assert orc.kept(l).tolist() == dbg[l]["kept"]
        if "blur" in dbg[l]:
            assert np.array_equal(orc.blurred(l), dbg[l]["blur"])
    rk = np.asarray(rk, np.float64)
    assert len(kps) == len(rk) >= nfeat
    assert np.array_equal(kps["x"], rk[:, 0].astype(np.float32)) and np.array_equal(kps["y"], rk[:, 1].astype(np.float32))
    assert np.array_equal(kps["size"], rk[:, 2].astype(np.float32)) and np.array_equal(kps["octave"], rk[:, 5].astype(np.int32))
    assert np.array_equal(kps["response"], rk[:, 4].astype(np.float32))
    assert np.abs(kps["angle"] - rk[:, 3]).max() <= 1e-3
    assert (desc == rd).all(1).mean() >= 0.999


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_oracle_against_golden_vectors(oracle, path):
    g = np.load(path)
    orc = oracle.OracleExtractor(int(g["nfeatures"]), 1.2, 8, 20, 7)
    kps, desc = orc(g["image"])
    assert len(kps) == len(g["kp_octave"])
    assert np.array_equal(np.stack([kps["x"], kps["y"]], 1).reshape(-1, 2), g["kp_xy"].reshape(-1, 2))
    assert np.array_equal(kps["octave"], g["kp_octave"]) and np.array_equal(kps["response"], g["kp_response"])
    assert np.array_equal(kps["size"], g["kp_size"])
    if len(kps):
        assert np.abs(kps["angle"] - g["kp_angle"]).max() <= 1e-3
        assert (desc == g["descriptors"]).all(1).mean() >= 0.999
    for l in range(8):
        assert zlib.crc32(orc.level(l).tobytes()) == int(g["level_crc"][l])
        b = orc.blurred(l)
        assert (0 if b is None else zlib.crc32(b.tobytes())) == int(g["blur_crc"][l])
        assert len(orc.candidates(l)[0]) == int(g["n_candidates"][l])
        assert len(orc.kept(l)) == int(g["n_kept"][l])
        assert orc.stats(l)["fallback"] == int(g["fallback_cells"][l])


def test_golden_set_is_present():
    assert len(glob.glob(os.path.join(GOLD, "*.npz"))) >= 6
