"""Pins the oracle's ordered matchers (SearchByProjection local-map / last-frame, SearchByBoW) against the
independent Python restatement in tests/search_reference.py (which itself uses cv2.gemm for the projection)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

import search_reference as R
from orb_slam2_detailed_comments_b200.synth import tracking_scene

SF = np.cumprod(np.concatenate([[1.0], np.full(7, np.float32(1.2), np.float32)]).astype(np.float32)).astype(np.float32)


def local_map_points(sc, q, seed):
    """MapPoint track fields (Frame::isInFrustum output) derived from projected queries."""
    rng = np.random.RandomState(seed)
    mps = []
    for i in range(len(q)):
        mps.append({"track_in_view": bool(q["flags"][i] & 1), "bad": rng.rand() < 0.03, "level": int(sc["last"]["octave"][i]),
                    "view_cos": np.float32(0.99 + 0.01 * rng.rand()), "x": q["u"][i], "y": q["v"][i], "xr": q["ur"][i],
                    "desc": sc["mp_desc"][i], "nobs": int((sc["mp_flags"][i] >> 1) & 1) * 3})
    return mps


def local_map_queries(oracle, mps, th):
    q = np.zeros(len(mps), oracle.PROJ_QUERY_DTYPE)
    for i, m in enumerate(mps):
        r = np.float32(2.5) if m["view_cos"] > np.float32(0.998) else np.float32(4.0)
        if th != 1.0:
            r = np.float32(r * np.float32(th))
        q[i] = (m["x"], m["y"], np.float32(r * SF[m["level"]]), m["xr"], 0, m["level"] - 1, m["level"],
                (1 if m["track_in_view"] and not m["bad"] else 0) | (2 if m["nobs"] > 0 else 0))
    return q


@pytest.mark.parametrize("seed,direction,th", [(1, 0, 15.0), (2, 1, 15.0), (3, 2, 7.0), (4, 0, 30.0)])
def test_last_frame_projection_search(oracle, seed, direction, th):
    sc = tracking_scene(600, 500, seed, frac_unobserved=0.2 if seed == 4 else 0.05)
    q = oracle.project_last_frame(sc["Xw"], sc["mp_flags"], sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], th, SF, direction)
    nm, mk, mq = oracle.search_by_projection(sc["cur"], sc["cur_desc"], sc["uright"], sc["bounds"], sc["occupied0"], q, sc["mp_desc"],
                                             oracle.SEARCH_BEST, 100, 0.0, True)
    F = R.PyFrame(sc["cur"], sc["cur_desc"], sc["bounds"], sc["uright"])
    nm_ref, mk_ref = R.search_last_frame(F, sc["occupied0"], sc["last"], sc["Xw"], sc["mp_flags"], sc["mp_desc"], sc["Tcw"], sc["cam4"],
                                         sc["mbf"], th, SF, direction)
    assert nm == nm_ref and nm > 50
    assert mk.tolist() == mk_ref


@pytest.mark.parametrize("seed,th", [(11, 1.0), (12, 3.0), (13, 5.0)])
def test_local_map_projection_search(oracle, seed, th):
    sc = tracking_scene(600, 700, seed, frac_mapped=0.9)
    q0 = oracle.project_last_frame(sc["Xw"], sc["mp_flags"] | 1, sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], 1.0, SF, 0)
    mps = local_map_points(sc, q0, seed)
    q = local_map_queries(oracle, mps, th)
    nm, mk, _ = oracle.search_by_projection(sc["cur"], sc["cur_desc"], sc["uright"], sc["bounds"], sc["occupied0"], q, sc["mp_desc"],
                                            oracle.SEARCH_RATIO_LEVEL, 100, 0.8, False)
    F = R.PyFrame(sc["cur"], sc["cur_desc"], sc["bounds"], sc["uright"])
    nm_ref, mk_ref = R.search_local_map(F, sc["occupied0"], mps, th, SF, 0.8)
    assert nm == nm_ref and nm > 30
    assert mk.tolist() == mk_ref


@pytest.mark.parametrize("seed,nodes,ori", [(21, 40, True), (22, 8, True), (23, 200, False)])
def test_search_by_bow(oracle, seed, nodes, ori):
    sc = tracking_scene(500, 450, seed, flip_bits=40)
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, nodes, 500).astype(np.int32)
    node1 = np.where(rng.rand(450) < 0.85, node2[sc["src"]], rng.randint(0, nodes, 450)).astype(np.int32)
    node1[rng.rand(450) < 0.02] = -1; node2[rng.rand(500) < 0.02] = -1
    usable = sc["mp_flags"] & 1
    nm, mk, mq = oracle.search_by_bow(sc["last"], sc["mp_desc"], node1, usable, sc["cur"], sc["cur_desc"], node2, 50, 0.7, ori)
    nm_ref, mk_ref = R.search_by_bow(sc["last"], sc["mp_desc"], node1, usable, sc["cur"], sc["cur_desc"], node2, 0.7, ori)
    assert nm == nm_ref and nm > 20
    assert mk.tolist() == mk_ref
    assert nm == int((mk >= 0).sum())        # every accepted match occupies its keypoint: no double assignment
    sel = np.nonzero(mq >= 0)[0]
    assert np.array_equal(mk[mq[sel]], sel)


def test_projection_arithmetic_matches_cv2_gemm(oracle):
    rng = np.random.RandomState(5)
    n = 5000
    Xw = (rng.randn(n, 3) * 20).astype(np.float32)
    T = np.eye(4, dtype=np.float32); T[:3, :4] = rng.randn(3, 4).astype(np.float32)
    k = np.zeros(n, oracle.KP_DTYPE)
    cam4 = np.array([500, 510, 320, 240], np.float32)
    big = np.array([-1e9, 1e9, -1e9, 1e9], np.float32)
    q = oracle.project_last_frame(Xw, np.ones(n, np.uint8), k, T, cam4, big, 40.0, 1.0, SF, 0)
    R3 = np.ascontiguousarray(T[:3, :3]); t = np.ascontiguousarray(T[:3, 3:4])
    for i in range(n):
        c = cv2.gemm(R3, Xw[i].reshape(3, 1), 1.0, t, 1.0)
        invz = np.float32(1.0 / np.float64(c[2, 0]))
        if invz < 0:
            assert q["flags"][i] == 0
            continue
        u = np.float32(np.float32(np.float32(cam4[0] * c[0, 0]) * invz) + cam4[2])
        v = np.float32(np.float32(np.float32(cam4[1] * c[1, 0]) * invz) + cam4[3])
        assert q["flags"][i] & 1 and q["u"][i] == u and q["v"][i] == v
        assert q["ur"][i] == np.float32(u - np.float32(np.float32(40.0) * invz))


def test_search_by_bow_keyframe_pair(oracle):
    """SearchByBoW(KeyFrame*, KeyFrame*) (ORBmatcher.cc:729): candidates need a map point, strict threshold."""
    sc = tracking_scene(500, 450, 31, flip_bits=60)
    rng = np.random.RandomState(31)
    node2 = rng.randint(0, 30, 500).astype(np.int32)
    node1 = np.where(rng.rand(450) < 0.85, node2[sc["src"]], rng.randint(0, 30, 450)).astype(np.int32)
    usable1 = sc["mp_flags"] & 1
    usable2 = (rng.rand(500) < 0.7).astype(np.uint8)
    nm, mk, mq = oracle.search_by_bow(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, 49, 0.8, True,
                                      unusable2=1 - usable2)
    nm_ref, mk_ref = R.search_by_bow(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, 0.8, True,
                                     usable2=usable2, strict=True)
    assert nm == nm_ref and nm > 20 and mk.tolist() == mk_ref
    assert np.all(usable2[mq[mq >= 0]] == 1)


SIGMA2 = (SF * SF).astype(np.float32)


@pytest.mark.parametrize("seed,only_stereo,mono", [(41, 0, False), (42, 1, False), (43, 0, True)])
def test_search_for_triangulation(oracle, seed, only_stereo, mono):
    from orb_slam2_detailed_comments_b200.synth import triangulation_pair
    sc = tracking_scene(500, 450, seed, flip_bits=50, noise_px=1.0)
    tp = triangulation_pair(sc, seed)
    rng = np.random.RandomState(seed)
    node2 = rng.randint(0, 25, 500).astype(np.int32)
    node1 = np.where(rng.rand(450) < 0.85, node2[sc["src"]], rng.randint(0, 25, 450)).astype(np.int32)
    ur1 = None if mono else tp["ur1"]; ur2 = None if mono else sc["uright"]
    pair = np.zeros(1, oracle.TRI_PAIR_DTYPE)
    pair["F12"][0] = tp["F12"]; pair["ex"], pair["ey"], pair["only_stereo"] = tp["ex"], tp["ey"], only_stereo
    nm, m12 = oracle.search_for_triangulation(tp["kps1"], sc["mp_desc"], node1, tp["has_mp1"], ur1, sc["cur"], sc["cur_desc"], node2,
                                              tp["has_mp2"], ur2, pair, SF, SIGMA2, True)
    nm_ref, m12_ref = R.search_for_triangulation(tp["kps1"], sc["mp_desc"], node1, tp["has_mp1"], ur1, sc["cur"], sc["cur_desc"], node2,
                                                 tp["has_mp2"], ur2, tp["F12"], tp["ex"], tp["ey"], only_stereo, SF, SIGMA2, True)
    assert nm == nm_ref and m12.tolist() == m12_ref
    assert nm > (5 if only_stereo else 15), nm
    sel = m12[m12 >= 0]
    assert len(set(sel.tolist())) == len(sel) and not tp["has_mp2"][sel].any() and not tp["has_mp1"][m12 >= 0].any()
