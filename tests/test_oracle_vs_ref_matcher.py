"""Pins the matcher / Frame half of the oracle against the REFERENCE ITSELF: src/ORBmatcher.cc and src/Frame.cc
compiled unmodified into oracle/_ref/liborbref.so (oracle/Makefile; OpenCV stand-in: oracle/cvshim/).
Covered: ORBmatcher::DescriptorDistance (:2083), ComputeThreeMaxima (:2035), SearchForInitialization (:573) with
Frame::GetFeaturesInArea (Frame.cc:590), Frame::UndistortKeyPoints (:724), ComputeImageBounds (:779),
AssignFeaturesToGrid (:399) and Frame::ComputeStereoMatches (:831). All comparisons are bit-exact."""
import numpy as np
import pytest

from orb_slam2_detailed_comments_b200.synth import correlated_descriptor_pair, random_descriptors, synth_frame

TUM1 = np.array([517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
EUROC = np.array([458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0], np.float32)
NODIST = np.array([718.856, 718.856, 607.1928, 185.2157, 0, 0, 0, 0, 0], np.float32)


def random_kps(dtype, n, w, h, seed):
    rng = np.random.RandomState(seed)
    k = np.zeros(n, dtype)
    k["x"] = (rng.rand(n) * w).astype(np.float32); k["y"] = (rng.rand(n) * h).astype(np.float32)
    k["octave"] = rng.randint(0, 8, n); k["angle"] = rng.rand(n) * 360; k["class_id"] = -1
    return k


def test_descriptor_distance_and_constants(oracle, reference):
    a = random_descriptors(300, 1); b = random_descriptors(300, 2)
    for i in range(300):
        assert reference.descriptor_distance(a[i], b[i]) == oracle.hamming(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())
    assert reference.descriptor_distance(a[0], a[0]) == 0 and reference.descriptor_distance(a[0], ~a[0]) == 256
    assert reference.matcher_constants() == dict(TH_LOW=50, TH_HIGH=100, HISTO_LENGTH=30)


def test_three_maxima(oracle, reference):
    rng = np.random.RandomState(0)
    for _ in range(500):
        h = rng.randint(0, rng.randint(1, 40), 30).astype(np.int32)
        if rng.rand() < 0.3:
            h[rng.randint(0, 30, 5)] = h.max()
        assert oracle.three_maxima(h) == reference.three_maxima(h)
    assert reference.three_maxima(np.zeros(30, np.int32)) == oracle.three_maxima(np.zeros(30, np.int32))


@pytest.mark.parametrize("cam,w,h", [(TUM1, 640, 480), (EUROC, 752, 480), (NODIST, 1241, 376)])
def test_frame_helpers(oracle, reference, cam, w, h):
    k = random_kps(oracle.KP_DTYPE, 3000, w, h, 2)
    d = random_descriptors(3000, 3)
    F = reference.ReferenceFrame(k, d, cam, w, h)
    un = oracle.undistort_keypoints(k, cam)
    assert F.keys_un().tobytes() == un.tobytes()                                  # UndistortKeyPoints
    b = oracle.image_bounds(cam, w, h)
    assert np.array_equal(F.bounds(), b)                                          # ComputeImageBounds
    start, items = oracle.assign_grid(un, b)
    rstart, ritems = F.grid()
    assert np.array_equal(start, rstart) and np.array_equal(items, ritems)       # AssignFeaturesToGrid
    rng = np.random.RandomState(3)
    n_hits = 0
    for _ in range(400):
        x, y, r = rng.rand() * (w + 60) - 30, rng.rand() * (h + 60) - 30, rng.rand() * 120 + 1
        lo, hi = [(-1, -1), (0, 0), (2, 5), (3, -1), (0, 7)][rng.randint(5)]
        got = oracle.features_in_area(un, b, x, y, r, lo, hi)
        ref = F.features_in_area(x, y, r, lo, hi)
        assert np.array_equal(got, ref)                                           # GetFeaturesInArea, order included
        n_hits += len(ref)
    assert n_hits > 1000


def _init_scene(n, seed, w=640, h=480, outlier_frac=0.2, sigma=15):
    rng = np.random.RandomState(seed)
    A, B, aa, ab = correlated_descriptor_pair(n, seed, outlier_frac=outlier_frac)
    xy1 = np.stack([rng.rand(n) * w, rng.rand(n) * h], 1).astype(np.float32)
    D = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(2)
    nn = D.argmin(0)
    xy2 = (xy1[nn] + rng.normal(0, sigma, (n, 2))).astype(np.float32)
    oc1 = (rng.rand(n) < 0.3).astype(np.int32) * rng.randint(1, 8, n)
    oc2 = (rng.rand(n) < 0.3).astype(np.int32) * rng.randint(1, 8, n)
    return A, B, aa, ab, xy1, xy2, oc1, oc2


def _kps(dtype, xy, octave, angle):
    k = np.zeros(len(xy), dtype)
    k["x"] = xy[:, 0]; k["y"] = xy[:, 1]; k["octave"] = octave; k["angle"] = angle; k["class_id"] = -1
    return k


@pytest.mark.parametrize("n,seed", [(400, 9), (1000, 11), (2000, 12)])
def test_search_for_initialization(oracle, reference, n, seed):
    w, h = 640, 480
    A, B, aa, ab, xy1, xy2, oc1, oc2 = _init_scene(n, seed, w, h)
    cam = np.array([500, 500, 320, 240, 0, 0, 0, 0, 0], np.float32)
    F1 = reference.ReferenceFrame(_kps(oracle.KP_DTYPE, xy1, oc1, aa), A, cam, w, h)
    F2 = reference.ReferenceFrame(_kps(oracle.KP_DTYPE, xy2, oc2, ab), B, cam, w, h)
    for window, ratio, ori in ((100, 0.9, True), (25, 0.9, True), (100, 0.9, False), (60, 0.75, True), (400, 0.9, True)):
        rn, rm12, rprev = reference.search_for_initialization(F1, F2, xy1, window, ratio, ori)
        on, om12, oprev, _, _ = oracle.search_for_initialization(xy1, oc1, aa, A, xy2, oc2, ab, B, (0, w, 0, h), xy1, window, ratio, ori, 0)
        assert rn == on and np.array_equal(rm12, om12), (window, ratio, ori)
        assert np.array_equal(rprev, oprev)
        assert rn > 10


def test_search_for_initialization_on_extracted_frames(oracle, reference):
    """Two views of one synthetic scene through the reference extractor, then the reference matcher, against the
    oracle's extractor + matcher: the monocular initialisation front-end end to end (Tracking.cc:915-926)."""
    w, h, nfeat = 640, 480, 1000
    big = synth_frame(w + 16, h + 16, 1, noise_sigma=0.0).astype(np.float32)
    rng = np.random.RandomState(0)
    a = np.clip(np.rint(big[:h, :w] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    b = np.clip(np.rint(big[4:h + 4, 7:w + 7] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    R = reference.ReferenceExtractor(nfeat, 1.2, 8, 20, 7)
    O = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    (ka, da), (kb, db) = R(a), R(b)
    (oka, oda), (okb, odb) = O(a), O(b)
    assert ka.tobytes() == oka.tobytes() and kb.tobytes() == okb.tobytes()
    F1 = reference.ReferenceFrame(ka, da, TUM1, w, h); F2 = reference.ReferenceFrame(kb, db, TUM1, w, h)
    u1 = oracle.undistort_keypoints(oka, TUM1); u2 = oracle.undistort_keypoints(okb, TUM1)
    bounds = oracle.image_bounds(TUM1, w, h)
    prev = np.stack([u1["x"], u1["y"]], 1)
    rn, rm12, rprev = reference.search_for_initialization(F1, F2, prev, 100, 0.9, True)
    on, om12, oprev, _, _ = oracle.search_for_initialization(prev, u1["octave"], u1["angle"], oda, np.stack([u2["x"], u2["y"]], 1),
                                                             u2["octave"], u2["angle"], odb, bounds, prev, 100, 0.9, True, 0)
    assert rn == on > 50 and np.array_equal(rm12, om12) and np.array_equal(rprev, oprev)


def stereo_pair(w, h, seed, disparity=11):
    big = synth_frame(w + 64, h, seed, noise_sigma=0).astype(np.float32)
    rng = np.random.RandomState(seed)
    left = np.clip(np.rint(big[:, 32:32 + w] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    right = np.clip(np.rint(big[:, 32 + disparity:32 + disparity + w] + rng.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    return left, right


@pytest.mark.parametrize("w,h,nfeat,disp,seed", [(640, 480, 800, 9, 3), (752, 480, 1200, 23, 4), (1241, 376, 2000, 30, 5)])
def test_compute_stereo_matches(oracle, reference, w, h, nfeat, disp, seed):
    left, right = stereo_pair(w, h, seed, disp)
    eL = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7); eR = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    rL = reference.ReferenceExtractor(nfeat, 1.2, 8, 20, 7); rR = reference.ReferenceExtractor(nfeat, 1.2, 8, 20, 7)
    kl, dl = eL(left); kr, dr = eR(right)
    rkl, rdl = rL(left); rkr, rdr = rR(right)
    assert kl.tobytes() == rkl.tobytes() and kr.tobytes() == rkr.tobytes()
    mbf, mb = 386.1448, 0.5371
    ur, dp, n = oracle.stereo_matches(eL, eR, kl, dl, kr, dr, mbf, mb)
    rur, rdp, rn = reference.stereo_matches(rL, rR, mbf, mb)
    assert np.array_equal(ur.view(np.uint32), rur.view(np.uint32))     # mvuRight, bit-exact
    assert np.array_equal(dp.view(np.uint32), rdp.view(np.uint32))     # mvDepth, bit-exact
    assert n == rn > len(kl) // 3


# ---- the projection searches of Tracking, on live ORB_SLAM2::MapPoint objects (src/MapPoint.cc, src/Map.cc) ----
from orb_slam2_detailed_comments_b200.synth import tracking_scene  # noqa: E402
from test_oracle_search import SF, local_map_points, local_map_queries  # noqa: E402


def _cam9(sc):
    return np.concatenate([sc["cam4"], np.zeros(5, np.float32)])


@pytest.mark.parametrize("seed,direction,th,n_cur,n_last", [(1, 0, 15.0, 600, 500), (2, 1, 15.0, 600, 500), (3, 2, 7.0, 600, 500),
                                                            (4, 0, 30.0, 600, 500), (5, 0, 15.0, 2000, 2000)])
def test_search_by_projection_last_frame(oracle, reference, seed, direction, th, n_cur, n_last):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (:1710) incl. its projection (:1734-1775)."""
    sc = tracking_scene(n_cur, n_last, seed, frac_unobserved=0.2 if seed == 4 else 0.05)
    q = oracle.project_last_frame(sc["Xw"], sc["mp_flags"], sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], th, SF, direction)
    nm, mk, mq = oracle.search_by_projection(sc["cur"], sc["cur_desc"], sc["uright"], sc["bounds"], sc["occupied0"], q, sc["mp_desc"],
                                             oracle.SEARCH_BEST, 100, 0.0, True)
    F = reference.ReferenceFrame(sc["cur"], sc["cur_desc"], _cam9(sc), 1241, 376)
    rn, rmk = reference.search_last_frame(F, sc["uright"], sc["occupied0"], sc["last"], sc["Xw"], sc["mp_flags"], sc["mp_desc"], sc["Tcw"],
                                          sc["cam4"], sc["mbf"], sc["mb"], th, direction, SF)
    assert rn == nm > 50
    assert np.array_equal(rmk, mk)


@pytest.mark.parametrize("seed,th", [(11, 1.0), (12, 3.0), (13, 5.0)])
def test_search_by_projection_local_map(oracle, reference, seed, th):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (:72), as Tracking::SearchLocalPoints calls it (nnratio 0.8)."""
    sc = tracking_scene(600, 700, seed, frac_mapped=0.9)
    q0 = oracle.project_last_frame(sc["Xw"], sc["mp_flags"] | 1, sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], 1.0, SF, 0)
    mps = local_map_points(sc, q0, seed)
    q = local_map_queries(oracle, mps, th)
    nm, mk, _ = oracle.search_by_projection(sc["cur"], sc["cur_desc"], sc["uright"], sc["bounds"], sc["occupied0"], q, sc["mp_desc"],
                                            oracle.SEARCH_RATIO_LEVEL, 100, 0.8, False)
    F = reference.ReferenceFrame(sc["cur"], sc["cur_desc"], _cam9(sc), 1241, 376)
    rn, rmk = reference.search_local_map(F, sc["uright"], sc["occupied0"], mps, th, 0.8, sc["cam4"], sc["mbf"], sc["mb"], SF)
    assert rn == nm > 30
    assert np.array_equal(rmk, mk)


# ---- the bag-of-words guided searches, on live ORB_SLAM2::KeyFrame objects (src/KeyFrame.cc) ----
def _nodes(sc, seed, n1, n2, nodes):
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, nodes, n2).astype(np.int32)
    node1 = np.where(rng.rand(n1) < 0.85, node2[sc["src"]], rng.randint(0, nodes, n1)).astype(np.int32)
    node1[rng.rand(n1) < 0.02] = -1; node2[rng.rand(n2) < 0.02] = -1
    return node1, node2


@pytest.mark.parametrize("seed,nodes,ori,n2,n1", [(21, 40, True, 500, 450), (22, 8, True, 500, 450), (23, 200, False, 500, 450),
                                                  (24, 60, True, 2000, 1800)])
def test_search_by_bow_keyframe_to_frame(oracle, reference, seed, nodes, ori, n2, n1):
    """ORBmatcher(0.7, true).SearchByBoW(pKF, F, vpMapPointMatches) (:247), as TrackReferenceKeyFrame calls it."""
    sc = tracking_scene(n2, n1, seed, flip_bits=40)
    node1, node2 = _nodes(sc, seed, n1, n2, nodes)
    usable = sc["mp_flags"] & 1
    nm, mk, mq = oracle.search_by_bow(sc["last"], sc["mp_desc"], node1, usable, sc["cur"], sc["cur_desc"], node2, 50, 0.7, ori)
    rn, rmk = reference.search_by_bow_frame(sc["last"], sc["mp_desc"], node1, usable, sc["cur"], sc["cur_desc"], node2, 0.7, ori, SF)
    assert rn == nm > 20
    assert np.array_equal(rmk, mk)


@pytest.mark.parametrize("seed", [31, 32])
def test_search_by_bow_keyframe_pair(oracle, reference, seed):
    """ORBmatcher(0.8, true).SearchByBoW(pKF1, pKF2, vpMatches12) (:729), as LoopClosing calls it."""
    sc = tracking_scene(500, 450, seed, flip_bits=60)
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, 30, 500).astype(np.int32)
    node1 = np.where(rng.rand(450) < 0.85, node2[sc["src"]], rng.randint(0, 30, 450)).astype(np.int32)
    usable1 = sc["mp_flags"] & 1
    usable2 = (rng.rand(500) < 0.7).astype(np.uint8)
    nm, mk, mq = oracle.search_by_bow(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, 49, 0.8, True,
                                      unusable2=1 - usable2)
    rn, rm12 = reference.search_by_bow_keyframes(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, usable2,
                                                 0.8, True, SF)
    assert rn == nm > 20
    assert np.array_equal(rm12, mq)


@pytest.mark.parametrize("seed,only_stereo,mono", [(41, 0, False), (42, 1, False), (43, 0, True), (44, 0, False)])
def test_search_for_triangulation(oracle, reference, seed, only_stereo, mono):
    """ORBmatcher(0.6, false/true).SearchForTriangulation (:884) incl. CheckDistEpipolarLine (:205) and the epipole (:898)."""
    from orb_slam2_detailed_comments_b200.synth import triangulation_pair
    sc = tracking_scene(500, 450, seed, flip_bits=50, noise_px=1.0)
    tp = triangulation_pair(sc, seed)
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, 25, 500).astype(np.int32)
    node1 = np.where(rng.rand(450) < 0.85, node2[sc["src"]], rng.randint(0, 25, 450)).astype(np.int32)
    ur1 = None if mono else tp["ur1"]; ur2 = None if mono else sc["uright"]
    sigma2 = (SF * SF).astype(np.float32)
    # the epipole exactly as :898-908 computes it for keyframe 1 at the identity: C2 = t2w, float arithmetic
    f32 = np.float32
    t = sc["Tcw"][:3, 3].astype(f32); fx, fy, cx, cy = [f32(v) for v in sc["cam4"]]
    invz = f32(1.0) / t[2]
    ex = f32(f32(f32(fx * t[0]) * invz) + cx); ey = f32(f32(f32(fy * t[1]) * invz) + cy)
    pair = np.zeros(1, oracle.TRI_PAIR_DTYPE)
    pair["F12"][0] = tp["F12"]; pair["ex"], pair["ey"], pair["only_stereo"] = ex, ey, only_stereo
    nm, m12 = oracle.search_for_triangulation(tp["kps1"], sc["mp_desc"], node1, tp["has_mp1"], ur1, sc["cur"], sc["cur_desc"], node2,
                                              tp["has_mp2"], ur2, pair, SF, sigma2, True)
    rn, rm12 = reference.search_for_triangulation(tp["kps1"], sc["mp_desc"], node1, tp["has_mp1"], ur1, sc["cur"], sc["cur_desc"], node2,
                                                  tp["has_mp2"], ur2, sc["Tcw"], sc["cam4"], tp["F12"], only_stereo, True, SF, sigma2)
    assert rn == nm > (5 if only_stereo else 15)
    assert np.array_equal(rm12, m12)


def test_compute_distinctive_descriptors(oracle, reference):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:365) on live MapPoint / KeyFrame objects: the descriptor the
    reference keeps is the one the oracle selects (ties between equal medians included)."""
    rng = np.random.RandomState(9)
    counts = [1, 2, 3, 4, 7, 16, 31, 32, 33, 64, 100, 5, 2, 9]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    desc = rng.randint(0, 256, (offsets[-1], 32)).astype(np.uint8)
    for p, c in enumerate(counts):
        if c > 2:   # clustered observations with duplicates: ties in the medians
            base = desc[offsets[p]].copy()
            for i in range(c):
                d = base.copy()
                bits = rng.randint(0, 256, rng.randint(0, 40))
                np.bitwise_xor.at(d, bits >> 3, (1 << (bits & 7)).astype(np.uint8))
                desc[offsets[p] + i] = d
            desc[offsets[p] + c - 1] = desc[offsets[p]]
    best = oracle.distinctive_descriptors(desc, offsets)
    kept = reference.distinctive_descriptors(desc, offsets)
    for p in range(len(counts)):
        assert np.array_equal(kept[p], desc[offsets[p] + best[p]]), p


def test_product_local_map_query_helper_feeds_the_reference_result(oracle, reference):
    """The product's host helper search.local_map_queries (the radius of ORBmatcher.cc:88-100 / RadiusByViewingCos
    :171-177 from the MapPoint track fields) builds the same queries as the test-side construction that is pinned against
    the reference above, and the oracle run on them reproduces the reference's SearchByProjection assignment."""
    from orb_slam2_detailed_comments_b200 import search
    for seed, th in ((14, 1.0), (15, 3.0), (16, 5.0)):
        sc = tracking_scene(700, 800, seed, frac_mapped=0.9)
        q0 = oracle.project_last_frame(sc["Xw"], sc["mp_flags"] | 1, sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], 1.0, SF, 0)
        mps = local_map_points(sc, q0, seed)
        q_test = local_map_queries(oracle, mps, th)
        q_prod = search.local_map_queries([m["x"] for m in mps], [m["y"] for m in mps], [m["xr"] for m in mps],
                                          [m["level"] for m in mps], [m["view_cos"] for m in mps],
                                          [int(m["track_in_view"] and not m["bad"]) for m in mps], [int(m["nobs"] > 0) for m in mps], th, SF)
        for f in ("u", "v", "radius", "ur", "min_level", "max_level", "flags"):
            assert np.array_equal(q_prod[f], q_test[f]), f
        nm, mk, _ = oracle.search_by_projection(sc["cur"], sc["cur_desc"], sc["uright"], sc["bounds"], sc["occupied0"], q_prod.view(oracle.PROJ_QUERY_DTYPE),
                                                sc["mp_desc"], oracle.SEARCH_RATIO_LEVEL, 100, 0.8, False)
        F = reference.ReferenceFrame(sc["cur"], sc["cur_desc"], _cam9(sc), 1241, 376)
        rn, rmk = reference.search_local_map(F, sc["uright"], sc["occupied0"], mps, th, 0.8, sc["cam4"], sc["mbf"], sc["mb"], SF)
        assert rn == nm > 30 and np.array_equal(rmk, mk)
