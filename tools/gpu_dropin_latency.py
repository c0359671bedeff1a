#!/usr/bin/env python
"""Per-call wall time of the reference's own ORBmatcher / Frame members with the reference's CPU bodies (liborbref.so) and
with the drop-in bodies (liborbref_gpu.so -> liborb_b200.so), on the same live Frame / KeyFrame / MapPoint objects.

The clock sits inside oracle/ref_wrap.cpp around the member call only (object construction is not timed); the same wrapper
is linked into both libraries. Every case is run REPS times on freshly built objects, the first run is a warm-up, the median
of the rest is printed. Usage (GPU box):  python tools/gpu_dropin_latency.py > gpurun_out/dropin_latency.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REPS = 7


def med(api, fn):
    t = []
    out = None
    for _ in range(REPS):
        out = fn()
        t.append(api.last_call_us())
    return float(np.median(t[1:])), out


def main():
    from oracle import orb_ref as ref, orb_refgpu, orb_oracle
    from orb_slam2_detailed_comments_b200.synth import tracking_scene, triangulation_pair, synth_frame
    from test_oracle_search import SF, local_map_points
    gref = orb_refgpu.reference_api()
    rows = []

    def case(name, make):
        r_us, r_out = med(ref, lambda: make(ref))
        g_us, g_out = med(gref, lambda: make(gref))
        same = all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(r_out, g_out))
        rows.append((name, r_us, g_us, same))
        print("%-78s reference CPU %9.1f us   drop-in %8.1f us   x%5.1f   identical: %s" % (name, r_us, g_us, r_us / g_us, same), flush=True)

    # tracking: SearchByProjection(CurrentFrame, LastFrame, th, mono) - once per tracked frame (Tracking.cc:1047)
    for n, q in ((1000, 1000), (2000, 2000)):
        sc = tracking_scene(n, q, 310 + n, w=1241, h=376, distinct=0.9)
        cam9 = np.concatenate([sc["cam4"], np.zeros(5, np.float32)])
        args = (sc["uright"], sc["occupied0"], sc["last"], sc["Xw"], sc["mp_flags"], sc["mp_desc"], sc["Tcw"], sc["cam4"], sc["mbf"], sc["mb"],
                15.0, 0, SF)
        case("SearchByProjection(CurrentFrame, LastFrame, 15)  %d keypoints x %d map points" % (n, q),
             lambda api: api.search_last_frame(api.ReferenceFrame(sc["cur"], sc["cur_desc"], cam9, 1241, 376), *args))
    # local map: SearchByProjection(F, vpMapPoints, th) (Tracking::SearchLocalPoints)
    for n, q in ((2000, 2200), (2000, 6000)):
        sc = tracking_scene(n, q, 21 + q, frac_mapped=0.9)
        q0 = orb_oracle.project_last_frame(sc["Xw"], sc["mp_flags"] | 1, sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], 1.0, SF, 0)
        mps = local_map_points(sc, q0, 5)
        cam9 = np.concatenate([sc["cam4"], np.zeros(5, np.float32)])
        args = (sc["uright"], sc["occupied0"], mps, 1.0, 0.8, sc["cam4"], sc["mbf"], sc["mb"], SF)
        case("SearchByProjection(F, vpMapPoints, 1)  %d keypoints x %d local map points" % (n, q),
             lambda api: api.search_local_map(api.ReferenceFrame(sc["cur"], sc["cur_desc"], cam9, 1241, 376), *args))
    # bag of words: keyframe -> frame, keyframe -> keyframe
    for n2, n1, nodes in ((2000, 1800, 60), (2000, 1800, 600)):
        sc = tracking_scene(n2, n1, 24, flip_bits=40)
        rng = np.random.RandomState(24)
        node2 = rng.randint(0, nodes, n2).astype(np.int32)
        node1 = np.where(rng.rand(n1) < 0.85, node2[sc["src"]], rng.randint(0, nodes, n1)).astype(np.int32)
        args = (sc["last"], sc["mp_desc"], node1, sc["mp_flags"] & 1, sc["cur"], sc["cur_desc"], node2, 0.7, True, SF)
        case("SearchByBoW(pKF, F)  %d x %d features, %d vocabulary nodes" % (n1, n2, nodes), lambda api: api.search_by_bow_frame(*args))
        usable2 = (rng.rand(n2) < 0.7).astype(np.uint8)
        args2 = (sc["last"], sc["mp_desc"], node1, sc["mp_flags"] & 1, sc["cur"], sc["cur_desc"], node2, usable2, 0.8, True, SF)
        case("SearchByBoW(pKF1, pKF2)  %d x %d features, %d vocabulary nodes" % (n1, n2, nodes), lambda api: api.search_by_bow_keyframes(*args2))
    # triangulation
    sc = tracking_scene(2000, 1900, 44, flip_bits=50, noise_px=1.0)
    tp = triangulation_pair(sc, 44)
    rng = np.random.RandomState(44)
    node2 = rng.randint(0, 60, 2000).astype(np.int32)
    node1 = np.where(rng.rand(1900) < 0.85, node2[sc["src"]], rng.randint(0, 60, 1900)).astype(np.int32)
    argsT = (tp["kps1"], sc["mp_desc"], node1, tp["has_mp1"], tp["ur1"], sc["cur"], sc["cur_desc"], node2, tp["has_mp2"], sc["uright"], sc["Tcw"],
             sc["cam4"], tp["F12"], 0, True, SF, (SF * SF).astype(np.float32))
    case("SearchForTriangulation(pKF1, pKF2, F12)  1900 x 2000 features, 60 nodes", lambda api: api.search_for_triangulation(*argsT))
    # monocular initialisation on real extractions
    w, h = 1241, 376
    big = synth_frame(w + 16, h + 16, 5, noise_sigma=0.0).astype(np.float32)
    rs = np.random.RandomState(1)
    a = np.clip(np.rint(big[:h, :w] + rs.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    b = np.clip(np.rint(big[3:h + 3, 6:w + 6] + rs.normal(0, 2, (h, w))), 0, 255).astype(np.uint8)
    ex = ref.ReferenceExtractor(2000, 1.2, 8, 20, 7)
    (k1, d1), (k2, d2) = ex(a), ex(b)
    cam9 = np.array([718.856, 718.856, 607.1928, 185.2157, 0, 0, 0, 0, 0], np.float32)

    def sfi(api):
        F1, F2 = api.ReferenceFrame(k1, d1, cam9, w, h), api.ReferenceFrame(k2, d2, cam9, w, h)
        prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
        return api.search_for_initialization(F1, F2, prev, 100, 0.9, True)[:2]
    case("SearchForInitialization(F1, F2, window 100)  %d x %d keypoints" % (len(k1), len(k2)), sfi)
    bad = [r[0] for r in rows if not r[3]]
    if bad:
        print("RESULTS DIFFER:", bad)
        sys.exit(1)


if __name__ == "__main__":
    main()
