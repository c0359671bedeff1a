"""GPU parity for the reference's own multi-extractor usage:

* src/Tracking.cc:175-188 keeps ORBextractor(nFeatures, ...) and ORBextractor(2 * nFeatures, ...) alive side by side and
  alternates between them (monocular initialisation, and again after every tracking reset, :387-405);
* src/Frame.cc:146-154 runs the left and the right extractor from two threads at once.

Kernel attributes such as the dynamic shared-memory limit belong to the FUNCTION, not to a handle: a second handle with a
smaller geometry must not lower what the first one needs. Every result is compared with the oracle.
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(gpu, orc, img, tag):
    kps, desc = gpu(img)
    okps, odesc = orc(img)
    assert len(kps) == len(okps), "%s: %d keypoints vs oracle %d" % (tag, len(kps), len(okps))
    for f in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(kps[f], okps[f]), "%s: keypoint field %s differs" % (tag, f)
    if len(kps):
        d = np.abs(kps["angle"] - okps["angle"])
        assert np.minimum(d, 360.0 - d).max() <= 1e-3, tag
        assert (desc == odesc).all(1).mean() >= 0.999, tag
    return len(kps)


def test_two_feature_budgets_interleaved_and_a_third_image_size(oracle):
    """Handles (2000, ...) and (1000, ...) interleaved A, B, A, B on 640x480 (Tracking.cc:175-188), a third handle on
    1241x376 in between, then a 'reset' (new handles, old ones destroyed) and the same again."""
    from orb_slam2_detailed_comments_b200 import ORBextractor
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    imgs = [synth_frame(640, 480, s) for s in (21, 22, 23, 24)]
    wide = [synth_frame(1241, 376, s) for s in (31, 32)]
    for round_ in range(2):   # Tracking::Reset re-creates nothing in the reference, but a System restart does
        ini = ORBextractor(2000, 1.2, 8, 20, 7)     # mpIniORBextractor
        left = ORBextractor(1000, 1.2, 8, 20, 7)    # mpORBextractorLeft
        kitti = ORBextractor(2000, 1.2, 8, 20, 7)
        o_ini = oracle.OracleExtractor(2000, 1.2, 8, 20, 7)
        o_left = oracle.OracleExtractor(1000, 1.2, 8, 20, 7)
        n = []
        for i, img in enumerate(imgs):
            n.append(_check(ini, o_ini, img, "round %d ini frame %d" % (round_, i)))
            n.append(_check(left, o_left, img, "round %d left frame %d" % (round_, i)))
            if i < 2:
                n.append(_check(kitti, o_ini, wide[i], "round %d kitti frame %d" % (round_, i)))
        # back to the big one after the small ones ran (the order that used to lower the shared-memory limit)
        n.append(_check(ini, o_ini, imgs[0], "round %d ini again" % round_))
        n.append(_check(kitti, o_ini, wide[0], "round %d kitti again" % round_))
        assert min(n) >= 1000
        for e in (ini, left, kitti):
            e.close()


def test_batch_and_single_handles_with_different_geometries(oracle):
    """A throughput handle (batch of 8 EuRoC frames) and a per-frame TUM1 handle used alternately."""
    import torch
    from orb_slam2_detailed_comments_b200 import ORBextractor
    from orb_slam2_detailed_comments_b200._lib import KP_DTYPE
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    big = ORBextractor(1200, 1.2, 8, 20, 7, max_batch=8)
    small = ORBextractor(500, 1.2, 8, 20, 7)
    o_big = oracle.OracleExtractor(1200, 1.2, 8, 20, 7)
    o_small = oracle.OracleExtractor(500, 1.2, 8, 20, 7)
    frames = np.stack([synth_frame(752, 480, 40 + i) for i in range(8)])
    tum = synth_frame(640, 480, 50)
    cap = big.max_keypoints_for(752, 480)
    for rep in range(2):
        d_imgs = torch.from_numpy(frames).cuda()
        d_kps = torch.zeros((8, cap, 28), dtype=torch.uint8, device="cuda")
        d_desc = torch.zeros((8, cap, 32), dtype=torch.uint8, device="cuda")
        d_counts = torch.zeros(8, dtype=torch.int32, device="cuda")
        big.extract_batch_device(d_imgs, d_kps, d_desc, d_counts)   # torch's current stream by default
        _check(small, o_small, tum, "small rep %d" % rep)
        big.synchronize()
        counts = d_counts.cpu().numpy()
        kps = d_kps.cpu().numpy().view(KP_DTYPE).reshape(8, cap)
        desc = d_desc.cpu().numpy()
        for b in (0, 3, 7):
            okps, odesc = o_big(frames[b])
            assert counts[b] == len(okps)
            for f in ("x", "y", "octave", "response"):
                assert np.array_equal(kps[b, :counts[b]][f], okps[f]), (b, f)
            assert (desc[b, :counts[b]] == odesc).all(1).mean() >= 0.999
    big.close(); small.close()


def test_two_extractors_from_two_threads(oracle):
    """Frame.cc:146-154: thread threadLeft(&Frame::ExtractORB, this, 0, imLeft); thread threadRight(...); join both.
    Two same-parameter handles, one per thread, many frames each; plus a third thread with another geometry."""
    from orb_slam2_detailed_comments_b200 import ORBextractor
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    specs = [(1241, 376, 2000, 60), (1241, 376, 2000, 70), (640, 480, 1000, 80)]
    exts = [ORBextractor(nf, 1.2, 8, 20, 7) for (_, _, nf, _) in specs]
    imgs = [[synth_frame(w, h, seed + i) for i in range(6)] for (w, h, _, seed) in specs]
    results = [[None] * 6 for _ in specs]
    errors = []

    def work(k):
        try:
            for rep in range(3):
                for i, img in enumerate(imgs[k]):
                    results[k][i] = exts[k](img)
        except Exception as e:  # noqa: BLE001 - reported below
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(specs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k, (w, h, nf, seed) in enumerate(specs):
        orc = oracle.OracleExtractor(nf, 1.2, 8, 20, 7)
        for i in (0, 5):
            kps, desc = results[k][i]
            okps, odesc = orc(imgs[k][i])
            assert len(kps) == len(okps)
            for f in ("x", "y", "octave", "response"):
                assert np.array_equal(kps[f], okps[f]), (k, i, f)
            assert (desc == odesc).all(1).mean() >= 0.999
    for e in exts:
        e.close()


def test_initialisation_extractor_feeds_the_matcher_at_4000_keypoints(oracle):
    """KITTI monocular: mpIniORBextractor = 2 * 2000 features (Tracking.cc:188) -> SearchForInitialization on ~4000
    keypoints per frame (Tracking.cc:915-926). ADVICE r01: frame 2 used to be capped at 3072 keypoints."""
    from orb_slam2_detailed_comments_b200 import FrameView, ORBextractor, ORBmatcher
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    w, h = 1241, 376
    big = synth_frame(w + 16, h + 16, 5, noise_sigma=0.0).astype(np.float32)
    rng = np.random.RandomState(1)
    a = np.clip(np.rint(big[:h, :w] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    b = np.clip(np.rint(big[3:h + 3, 6:w + 6] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    ex = ORBextractor(4000, 1.2, 8, 20, 7)
    (k1, d1), (k2, d2) = ex(a), ex(b)
    assert len(k1) > 3072 and len(k2) > 3072, (len(k1), len(k2))
    F1, F2 = FrameView.from_keypoints(k1, d1, w, h), FrameView.from_keypoints(k2, d2, w, h)
    m = ORBmatcher(0.9, True, max_keypoints=8192)
    for mode in (0, 1):
        prev = F1.xy.copy()
        oprev = F1.xy.copy()
        n_ref, m_ref, prev_ref, best_ref, second_ref = oracle.search_for_initialization(
            F1.xy, F1.octave, F1.angle, F1.descriptors, F2.xy, F2.octave, F2.angle, F2.descriptors, (0, w, 0, h), oprev,
            window=100, nnratio=0.9, check_ori=True, mode=mode)
        n, m12, best, second = m.SearchForInitialization(F1, F2, prev, 100, mode=mode, want_distances=True)
        assert n == n_ref and np.array_equal(m12, m_ref), "mode %d: %d vs %d matches" % (mode, n, n_ref)
        assert np.array_equal(best, best_ref) and np.array_equal(second, second_ref)
        if mode == 0:
            assert np.array_equal(prev, prev_ref)
            # the form a drop-in caller gets (no per-row distances): k_sfi_scan + k_sfi_walk at 4000 x 4000
            prev2 = F1.xy.copy()
            n2, m2 = m.SearchForInitialization(F1, F2, prev2, 100, mode=0)
            assert n2 == n_ref and np.array_equal(m2, m_ref) and np.array_equal(prev2, prev_ref)
        assert n > 100
    ex.close(); m.close()
