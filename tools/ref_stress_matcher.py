"""CPU stress of the matcher pinning: oracle vs the reference's own ORBmatcher.cc / Frame.cc on many random scenes.
usage: python tools/ref_stress_matcher.py [scenes]"""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import orb_oracle as O, orb_ref as R
from orb_slam2_detailed_comments_b200.synth import tracking_scene, triangulation_pair
from test_oracle_vs_ref_matcher import _init_scene, _kps, _cam9, _nodes
from test_oracle_search import SF, local_map_points, local_map_queries

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = {"init": 0, "last": 0, "local": 0, "bow": 0, "bowkf": 0, "tri": 0}
cnt = dict.fromkeys(bad, 0)
t0 = time.time()
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
for s in range(n):
    seed = 10000 + s + 100000 * (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    # SearchForInitialization
    nk = int(rng.choice([200, 500, 1000, 2000])); w, h = 640, 480
    A, B, aa, ab, xy1, xy2, oc1, oc2 = _init_scene(nk, seed, w, h, outlier_frac=float(rng.rand() * 0.4), sigma=float(3 + rng.rand() * 25))
    cam = np.array([500, 500, 320, 240, 0, 0, 0, 0, 0], np.float32)
    F1 = R.ReferenceFrame(_kps(O.KP_DTYPE, xy1, oc1, aa), A, cam, w, h); F2 = R.ReferenceFrame(_kps(O.KP_DTYPE, xy2, oc2, ab), B, cam, w, h)
    window = int(rng.choice([10, 30, 100, 300])); ratio = float(rng.choice([0.9, 0.75, 0.6])); ori = bool(rng.randint(2))
    rn, rm, rp = R.search_for_initialization(F1, F2, xy1, window, ratio, ori)
    on, om, op, _, _ = O.search_for_initialization(xy1, oc1, aa, A, xy2, oc2, ab, B, (0, w, 0, h), xy1, window, ratio, ori, 0)
    cnt["init"] += rn
    bad["init"] += not (rn == on and np.array_equal(rm, om) and np.array_equal(rp, op))
    # tracking searches
    nc, nlast = int(rng.choice([300, 800, 2000])), int(rng.choice([300, 900, 2000]))
    sc = tracking_scene(nc, nlast, seed, frac_unobserved=float(rng.rand() * 0.3), frac_mapped=0.6 + 0.4 * rng.rand())
    th = float(rng.choice([7.0, 15.0, 30.0])); direction = int(rng.randint(3))
    q = O.project_last_frame(sc["Xw"], sc["mp_flags"], sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], th, SF, direction)
    nm, mk, _ = O.search_by_projection(sc["cur"], sc["cur_desc"], sc["uright"], sc["bounds"], sc["occupied0"], q, sc["mp_desc"], O.SEARCH_BEST, 100, 0.0, True)
    F = R.ReferenceFrame(sc["cur"], sc["cur_desc"], _cam9(sc), 1241, 376)
    rn, rmk = R.search_last_frame(F, sc["uright"], sc["occupied0"], sc["last"], sc["Xw"], sc["mp_flags"], sc["mp_desc"], sc["Tcw"], sc["cam4"], sc["mbf"], sc["mb"], th, direction, SF)
    cnt["last"] += rn; bad["last"] += not (rn == nm and np.array_equal(rmk, mk))
    thl = float(rng.choice([1.0, 3.0, 5.0]))
    q0 = O.project_last_frame(sc["Xw"], sc["mp_flags"] | 1, sc["last"], sc["Tcw"], sc["cam4"], sc["bounds"], sc["mbf"], 1.0, SF, 0)
    mps = local_map_points(sc, q0, seed); ql = local_map_queries(O, mps, thl)
    nm, mk, _ = O.search_by_projection(sc["cur"], sc["cur_desc"], sc["uright"], sc["bounds"], sc["occupied0"], ql, sc["mp_desc"], O.SEARCH_RATIO_LEVEL, 100, 0.8, False)
    rn, rmk = R.search_local_map(F, sc["uright"], sc["occupied0"], mps, thl, 0.8, sc["cam4"], sc["mbf"], sc["mb"], SF)
    cnt["local"] += rn; bad["local"] += not (rn == nm and np.array_equal(rmk, mk))
    # BoW searches
    nodes = int(rng.choice([8, 40, 200]))
    node1, node2 = _nodes(sc, seed, nlast, nc, nodes)
    usable1 = sc["mp_flags"] & 1
    nm, mk, mq = O.search_by_bow(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, 50, 0.7, ori)
    rn, rmk = R.search_by_bow_frame(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, 0.7, ori, SF)
    cnt["bow"] += rn; bad["bow"] += not (rn == nm and np.array_equal(rmk, mk))
    usable2 = (rng.rand(nc) < 0.7).astype(np.uint8)
    nm, mk, mq = O.search_by_bow(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, 49, 0.8, True, unusable2=1 - usable2)
    rn, rm12 = R.search_by_bow_keyframes(sc["last"], sc["mp_desc"], node1, usable1, sc["cur"], sc["cur_desc"], node2, usable2, 0.8, True, SF)
    cnt["bowkf"] += rn; bad["bowkf"] += not (rn == nm and np.array_equal(rm12, mq))
    # triangulation
    sct = tracking_scene(500, 450, seed, flip_bits=50, noise_px=1.0)
    tp = triangulation_pair(sct, seed)
    n1t, n2t = _nodes(sct, seed + 1, 450, 500, 25)
    mono = bool(rng.randint(2)); only_stereo = 0 if mono else int(rng.randint(2))
    ur1 = None if mono else tp["ur1"]; ur2 = None if mono else sct["uright"]
    f32 = np.float32
    t = sct["Tcw"][:3, 3].astype(f32); fx, fy, cx, cy = [f32(v) for v in sct["cam4"]]
    invz = f32(1.0) / t[2]
    pair = np.zeros(1, O.TRI_PAIR_DTYPE)
    pair["F12"][0] = tp["F12"]; pair["ex"] = f32(f32(f32(fx * t[0]) * invz) + cx); pair["ey"] = f32(f32(f32(fy * t[1]) * invz) + cy)
    pair["only_stereo"] = only_stereo
    sig2 = (SF * SF).astype(np.float32)
    nm, m12 = O.search_for_triangulation(tp["kps1"], sct["mp_desc"], n1t, tp["has_mp1"], ur1, sct["cur"], sct["cur_desc"], n2t, tp["has_mp2"], ur2, pair, SF, sig2, True)
    rn, rm12 = R.search_for_triangulation(tp["kps1"], sct["mp_desc"], n1t, tp["has_mp1"], ur1, sct["cur"], sct["cur_desc"], n2t, tp["has_mp2"], ur2, sct["Tcw"], sct["cam4"], tp["F12"], only_stereo, True, SF, sig2)
    cnt["tri"] += rn; bad["tri"] += not (rn == nm and np.array_equal(rm12, m12))
print("%d scenes, matches per function %s, mismatching scenes %s, %.1f s" % (n, cnt, bad, time.time() - t0))
sys.exit(1 if any(bad.values()) else 0)
