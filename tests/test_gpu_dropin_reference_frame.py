"""The boundary, for real: the reference's UNMODIFIED src/Frame.cc and src/ORBmatcher.cc, compiled against the reference's
unmodified headers and linked against the drop-in (compat/orb_b200_extractor.cpp, orb_b200_matcher.cpp, orb_b200_frame.cpp
over liborb_b200.so) instead of src/ORBextractor.cc - oracle/_ref/liborbref_gpu.so, `make -C oracle refgpu`.

The reference's real Frame constructors run (src/Frame.cc:313-350 monocular, :121-158 stereo with its two extractor threads
and ComputeStereoMatches); every member they fill is compared with what the all-CPU reference (oracle/_ref/liborbref.so, the
reference's own ORBextractor.cc / Frame.cc, canonical heap order) computes for the same images.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TUM1 = (517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314)   # TUM1.yaml
KITTI = (718.856, 718.856, 607.1928, 185.2157, 0.0, 0.0, 0.0, 0.0, 0.0)                                       # KITTI00-02.yaml
KITTI_BF = 386.1448


@pytest.fixture(scope="module")
def dropin():
    from oracle import orb_refgpu
    if not orb_refgpu.available():
        orb_refgpu.build()
    if not orb_refgpu.available():
        pytest.skip("oracle/_ref/liborbref_gpu.so not built (needs /root/reference)")
    orb_refgpu.lib()
    return orb_refgpu


def _same_keys(a, b, tag):
    assert len(a) == len(b), "%s: %d keypoints vs reference %d" % (tag, len(a), len(b))
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(a[f], b[f]), "%s: keypoint field %s differs" % (tag, f)
    if len(a):
        d = np.abs(a["angle"] - b["angle"])
        assert np.minimum(d, 360.0 - d).max() <= 1e-3, tag


def _same_desc(a, b, tag):
    assert a.shape == b.shape, tag
    if len(a):
        assert (a == b).all(1).mean() >= 0.999, "%s: %d descriptor rows differ" % (tag, int((~(a == b).all(1)).sum()))


def test_monocular_frame_constructor(dropin, reference):
    """Frame(imGray, ts, extractor, voc, K, distCoef, bf, thDepth): ExtractORB -> UndistortKeyPoints -> AssignFeaturesToGrid."""
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    ex = dropin.DropInExtractor(1000, 1.2, 8, 20, 7)
    rex = reference.ReferenceExtractor(1000, 1.2, 8, 20, 7)
    for seed in (3, 4):
        img = synth_frame(640, 480, seed)
        F = dropin.DropInFrame.mono(ex, img, TUM1)
        rk, rd = rex(img, canonical=True)
        RF = reference.ReferenceFrame(rk, rd, TUM1, 640, 480)
        assert F.n == len(rk) >= 1000
        _same_keys(F.keys(0), rk, "mvKeys seed %d" % seed)
        _same_desc(F.descriptors(), rd, "mDescriptors seed %d" % seed)
        ku, rku = F.keys(1), RF.keys_un()
        assert np.array_equal(ku["x"], rku["x"]) and np.array_equal(ku["y"], rku["y"]), "mvKeysUn differs"
        assert np.array_equal(F.bounds(), RF.bounds())
        (s, it), (rs, rit) = F.grid(), RF.grid()
        assert np.array_equal(s, rs) and np.array_equal(it, rit), "mGrid differs"
        ur, dp = F.stereo_vectors()
        assert (ur == -1).all() and (dp == -1).all()   # :343-344
        sc = F.scale_tables()
        for a, b in zip(sc, (rex.scale, rex.inv_scale, rex.sigma2, rex.inv_sigma2)):
            assert np.array_equal(a, b)
        for l in (0, 3, 7):
            assert np.array_equal(ex.level(l), rex.level(l)), "mvImagePyramid[%d] differs" % l
        F.close()
    ex.close()


def test_stereo_frame_constructor(dropin, reference):
    """Frame(imLeft, imRight, ...): two ExtractORB threads (:146-154), ComputeStereoMatches (:157), undistortion, grid."""
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    w, h = 1241, 376
    exL, exR = dropin.DropInExtractor(2000, 1.2, 8, 20, 7), dropin.DropInExtractor(2000, 1.2, 8, 20, 7)
    rL, rR = reference.ReferenceExtractor(2000, 1.2, 8, 20, 7), reference.ReferenceExtractor(2000, 1.2, 8, 20, 7)
    mb = KITTI_BF / KITTI[0]
    for seed, disp in ((11, 9), (12, 23)):
        big = synth_frame(w + 64, h, seed)
        left, right = np.ascontiguousarray(big[:, disp:disp + w]), np.ascontiguousarray(big[:, :w])
        F = dropin.DropInFrame.stereo(exL, exR, left, right, KITTI, KITTI_BF)
        kl, dl = rL(left, canonical=True)
        kr, dr = rR(right, canonical=True)
        rur, rdp, kept = reference.stereo_matches(rL, rR, KITTI_BF, mb)
        _same_keys(F.keys(0), kl, "mvKeys")
        _same_keys(F.keys(2), kr, "mvKeysRight")
        _same_desc(F.descriptors(False), dl, "mDescriptors")
        _same_desc(F.descriptors(True), dr, "mDescriptorsRight")
        ur, dp = F.stereo_vectors()
        assert np.array_equal(ur, rur), "mvuRight differs in %d entries" % int((ur != rur).sum())
        assert np.array_equal(dp, rdp), "mvDepth differs"
        assert kept > 100 and int((dp > 0).sum()) == kept
        RF = reference.ReferenceFrame(kl, dl, KITTI, w, h)
        (s, it), (rs, rit) = F.grid(), RF.grid()
        assert np.array_equal(s, rs) and np.array_equal(it, rit), "mGrid differs"
        F.close()
    exL.close(); exR.close()


def test_monocular_initialisation_sequence(dropin, reference):
    """src/Tracking.cc:175-188 + :895-926: mpIniORBextractor (2 * nFeatures) makes two frames, ORBmatcher(0.9, true)
    .SearchForInitialization(mInitialFrame, mCurrentFrame, mvbPrevMatched, mvIniMatches, 100); the normal extractor stays
    alive beside it and is used in between."""
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    w, h = 640, 480
    ini, left = dropin.DropInExtractor(2000, 1.2, 8, 20, 7), dropin.DropInExtractor(1000, 1.2, 8, 20, 7)
    rini = reference.ReferenceExtractor(2000, 1.2, 8, 20, 7)
    big = synth_frame(w + 16, h + 16, 5, noise_sigma=0.0).astype(np.float32)
    rng = np.random.RandomState(2)
    a = np.clip(np.rint(big[:h, :w] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    b = np.clip(np.rint(big[5:h + 5, 8:w + 8] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    F1 = dropin.DropInFrame.mono(ini, a, TUM1)
    left(a)                                            # the other live extractor, smaller geometry, in between
    F2 = dropin.DropInFrame.mono(ini, b, TUM1)
    k1, d1 = rini(a, canonical=True)
    k2, d2 = rini(b, canonical=True)
    R1, R2 = reference.ReferenceFrame(k1, d1, TUM1, w, h), reference.ReferenceFrame(k2, d2, TUM1, w, h)
    prev = np.stack([R1.keys_un()["x"], R1.keys_un()["y"]], 1)
    n_ref, m_ref, prev_ref = reference.search_for_initialization(R1, R2, prev, 100, 0.9, True)
    n, m12, prev_out = dropin.search_for_initialization(F1, F2, prev, 100, 0.9, True)
    assert n == n_ref and np.array_equal(m12, m_ref) and np.array_equal(prev_out, prev_ref), (n, n_ref)
    assert n > 100
    assert dropin.descriptor_distance(d1[0], d2[0]) == reference.descriptor_distance(d1[0], d2[0])
    for F in (F1, F2):
        F.close()
    ini.close(); left.close()


def _tracking_scenes(seed):
    from orb_slam2_detailed_comments_b200.synth import tracking_scene
    return [tracking_scene(n, q, seed + i, w=1241, h=376, distinct=0.9) for i, (n, q) in enumerate(((1500, 1400), (2000, 2000), (400, 900)))]


@pytest.mark.parametrize("seed,th,direction", [(310, 15.0, 0), (410, 7.0, 1), (510, 15.0, 2)])
def test_tracking_search_through_the_reference_class(dropin, reference, seed, th, direction):
    """ORBmatcher(0.9, true).SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1710-1860, called once per
    tracked frame at src/Tracking.cc:1047) on live ORB_SLAM2::Frame / MapPoint objects: in liborbref_gpu.so the method's body is
    the drop-in (compat/orb_b200_matcher.cpp -> orb_search_by_projection_last_frame: projection, grid, scan and in-order
    commit on the GPU), in liborbref.so the reference's own CPU body. CurrentFrame.mvpMapPoints and the return value must be
    identical."""
    gref = dropin.reference_api()
    sf = np.cumprod(np.concatenate([[1.0], np.full(7, np.float32(1.2), np.float32)]).astype(np.float32)).astype(np.float32)
    total = 0
    for s in _tracking_scenes(seed):
        cam9 = np.concatenate([s["cam4"], np.zeros(5, np.float32)])
        args = (s["uright"], s["occupied0"], s["last"], s["Xw"], s["mp_flags"], s["mp_desc"], s["Tcw"], s["cam4"], s["mbf"], s["mb"], th,
                direction, sf)
        Fc = reference.ReferenceFrame(s["cur"], s["cur_desc"], cam9, 1241, 376)
        rn, rmk = reference.search_last_frame(Fc, *args)
        Fg = gref.ReferenceFrame(s["cur"], s["cur_desc"], cam9, 1241, 376)
        gn, gmk = gref.search_last_frame(Fg, *args)
        assert gn == rn, "return value differs: %d vs %d" % (gn, rn)
        assert np.array_equal(gmk, rmk), "CurrentFrame.mvpMapPoints differs in %d entries" % int((gmk != rmk).sum())
        total += rn
    print("seed", seed, "matches", total)
    assert total > 1000


@pytest.mark.parametrize("seed,th", [(21, 1.0), (22, 3.0), (23, 5.0)])
def test_local_map_search_through_the_reference_class(dropin, reference, oracle, seed, th):
    """ORBmatcher(0.8).SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:72-169, Tracking::SearchLocalPoints) on live
    MapPoint objects whose track fields were filled as Frame::isInFrustum does: drop-in body (orb_search_by_projection_host)
    against the reference's own CPU body - F.mvpMapPoints and the return value identical."""
    from orb_slam2_detailed_comments_b200.synth import tracking_scene
    from test_oracle_search import SF, local_map_points
    gref = dropin.reference_api()
    total = 0
    for k, (n_cur, n_mp) in enumerate(((600, 700), (2000, 2200))):
        sc = tracking_scene(n_cur, n_mp, seed + 100 * k, frac_mapped=0.9)
        q0 = oracle.project_last_frame(sc["Xw"], sc["mp_flags"] | 1, sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], 1.0, SF, 0)
        mps = local_map_points(sc, q0, seed)
        cam9 = np.concatenate([sc["cam4"], np.zeros(5, np.float32)])
        args = (sc["uright"], sc["occupied0"], mps, th, 0.8, sc["cam4"], sc["mbf"], sc["mb"], SF)
        rn, rmk = reference.search_local_map(reference.ReferenceFrame(sc["cur"], sc["cur_desc"], cam9, 1241, 376), *args)
        gn, gmk = gref.search_local_map(gref.ReferenceFrame(sc["cur"], sc["cur_desc"], cam9, 1241, 376), *args)
        assert gn == rn, "return value differs: %d vs %d" % (gn, rn)
        assert np.array_equal(gmk, rmk), "F.mvpMapPoints differs in %d entries" % int((gmk != rmk).sum())
        total += rn
    assert total > 100


def _bow_nodes(sc, seed, n1, n2, nodes):
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, nodes, n2).astype(np.int32)
    node1 = np.where(rng.rand(n1) < 0.85, node2[sc["src"]], rng.randint(0, nodes, n1)).astype(np.int32)
    node1[rng.rand(n1) < 0.02] = -1; node2[rng.rand(n2) < 0.02] = -1
    return node1, node2


@pytest.mark.parametrize("seed,nodes,ori,n2,n1", [(21, 40, True, 500, 450), (22, 8, True, 500, 450), (23, 200, False, 500, 450),
                                                  (24, 60, True, 2000, 1800)])
def test_bow_search_keyframe_to_frame_through_the_reference_class(dropin, reference, seed, nodes, ori, n2, n1):
    """ORBmatcher(0.7, true).SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:247-420, Tracking::TrackReferenceKeyFrame)
    on a live ORB_SLAM2::KeyFrame with live MapPoints: drop-in body (orb_search_by_bow_host) against the reference's CPU body."""
    from orb_slam2_detailed_comments_b200.synth import tracking_scene
    from test_oracle_search import SF
    gref = dropin.reference_api()
    sc = tracking_scene(n2, n1, seed, flip_bits=40)
    node1, node2 = _bow_nodes(sc, seed, n1, n2, nodes)
    args = (sc["last"], sc["mp_desc"], node1, sc["mp_flags"] & 1, sc["cur"], sc["cur_desc"], node2, 0.7, ori, SF)
    rn, rmk = reference.search_by_bow_frame(*args)
    gn, gmk = gref.search_by_bow_frame(*args)
    assert gn == rn > 20
    assert np.array_equal(gmk, rmk), "vpMapPointMatches differs in %d entries" % int((gmk != rmk).sum())


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_bow_search_keyframe_pair_through_the_reference_class(dropin, reference, seed):
    """ORBmatcher(0.8, true).SearchByBoW(pKF1, pKF2, vpMatches12) (src/ORBmatcher.cc:729-880, LoopClosing::ComputeSim3)."""
    from orb_slam2_detailed_comments_b200.synth import tracking_scene
    from test_oracle_search import SF
    gref = dropin.reference_api()
    n2, n1 = (500, 450) if seed < 33 else (1800, 1700)
    sc = tracking_scene(n2, n1, seed, flip_bits=60)
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, 30, n2).astype(np.int32)
    node1 = np.where(rng.rand(n1) < 0.85, node2[sc["src"]], rng.randint(0, 30, n1)).astype(np.int32)
    usable2 = (rng.rand(n2) < 0.7).astype(np.uint8)
    args = (sc["last"], sc["mp_desc"], node1, sc["mp_flags"] & 1, sc["cur"], sc["cur_desc"], node2, usable2, 0.8, True, SF)
    rn, rm12 = reference.search_by_bow_keyframes(*args)
    gn, gm12 = gref.search_by_bow_keyframes(*args)
    assert gn == rn > 20
    assert np.array_equal(gm12, rm12), "vpMatches12 differs in %d entries" % int((gm12 != rm12).sum())


@pytest.mark.parametrize("seed,only_stereo,mono,n2,n1", [(41, 0, False, 500, 450), (42, 1, False, 500, 450), (43, 0, True, 500, 450),
                                                         (44, 0, False, 2000, 1900)])
def test_triangulation_search_through_the_reference_class(dropin, reference, seed, only_stereo, mono, n2, n1):
    """ORBmatcher(0.6, true).SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (src/ORBmatcher.cc:884-1100,
    LocalMapping::CreateNewMapPoints): the epipole from the live keyframes' poses, CheckDistEpipolarLine on the GPU."""
    from orb_slam2_detailed_comments_b200.synth import tracking_scene, triangulation_pair
    from test_oracle_search import SF
    gref = dropin.reference_api()
    sc = tracking_scene(n2, n1, seed, flip_bits=50, noise_px=1.0)
    tp = triangulation_pair(sc, seed)
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, 25, n2).astype(np.int32)
    node1 = np.where(rng.rand(n1) < 0.85, node2[sc["src"]], rng.randint(0, 25, n1)).astype(np.int32)
    ur1 = None if mono else tp["ur1"]; ur2 = None if mono else sc["uright"]
    sigma2 = (SF * SF).astype(np.float32)
    args = (tp["kps1"], sc["mp_desc"], node1, tp["has_mp1"], ur1, sc["cur"], sc["cur_desc"], node2, tp["has_mp2"], ur2, sc["Tcw"],
            sc["cam4"], tp["F12"], only_stereo, True, SF, sigma2)
    rn, rm12 = reference.search_for_triangulation(*args)
    gn, gm12 = gref.search_for_triangulation(*args)
    assert gn == rn > (5 if only_stereo else 15)
    assert np.array_equal(gm12, rm12), "vMatchedPairs differs in %d entries" % int((gm12 != rm12).sum())


def test_stereo_and_initialisation_without_the_host_pyramid_copy():
    """ORB_B200_COMPAT_PYRAMID=0 (read once per process): the drop-in extractor leaves mvImagePyramid empty; the reference's stereo
    Frame constructor (ComputeStereoMatches is the drop-in, on the device-resident pyramid) and the initialisation sequence must
    still equal the all-CPU reference. Run in a child process because the switch is latched at the first call."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ORB_B200_COMPAT_PYRAMID="0")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                          "stereo_frame_constructor or monocular_initialisation_sequence"], env=env, capture_output=True, text=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "2 passed" in out.stdout, out.stdout[-500:]
