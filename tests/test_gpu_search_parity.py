"""GPU parity of the ordered matchers (SearchByProjection local map / last frame, SearchByBoW) through the C ABI
against the oracle: identical CurrentFrame.mvpMapPoints assignment (as indices), identical match count.
Batches are ragged (different N per frame, an empty frame, a frame without queries)."""
import numpy as np
import pytest

from orb_slam2_detailed_comments_b200.synth import tracking_scene

pytestmark = pytest.mark.gpu

SF = np.cumprod(np.concatenate([[1.0], np.full(7, np.float32(1.2), np.float32)]).astype(np.float32)).astype(np.float32)


class Batch:
    """Scenes packed into capacity-strided device tensors + the frame grid."""

    def __init__(self, scenes, cap, qcap):
        import torch
        from orb_slam2_detailed_comments_b200 import KP_DTYPE, frame, search
        self.scenes, self.cap, self.qcap = scenes, cap, qcap
        B = len(scenes)
        kps = np.zeros((B, cap), KP_DTYPE); desc = np.zeros((B, cap, 32), np.uint8); ur = np.full((B, cap), -1, np.float32)
        occ = np.zeros((B, cap), np.uint8); counts = np.zeros(B, np.int32)
        last = np.zeros((B, qcap), KP_DTYPE); Xw = np.zeros((B, qcap, 3), np.float32); fl = np.zeros((B, qcap), np.uint8)
        mpd = np.zeros((B, qcap, 32), np.uint8); qcounts = np.zeros(B, np.int32); T = np.zeros((B, 4, 4), np.float32)
        for b, s in enumerate(scenes):
            n, m = len(s["cur"]), len(s["last"])
            counts[b], qcounts[b] = n, m
            kps[b, :n], desc[b, :n], ur[b, :n], occ[b, :n] = s["cur"], s["cur_desc"], s["uright"], s["occupied0"]
            last[b, :m], Xw[b, :m], fl[b, :m], mpd[b, :m] = s["last"], s["Xw"], s["mp_flags"], s["mp_desc"]
            T[b] = s["Tcw"]
        self.bounds = scenes[0]["bounds"]
        dev = "cuda"
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.d_kps = t(kps.view(np.uint8).reshape(B, cap, 28)); self.d_desc = t(desc); self.d_ur = t(ur); self.d_occ = t(occ)
        self.d_counts = t(counts); self.d_last = t(last.view(np.uint8).reshape(B, qcap, 28)); self.d_Xw = t(Xw); self.d_fl = t(fl)
        self.d_mpd = t(mpd); self.d_qcounts = t(qcounts); self.d_T = t(T)
        self.d_cs = torch.zeros((B, 64 * 48 + 1), dtype=torch.int32, device=dev)
        self.d_ci = torch.zeros((B, cap), dtype=torch.int32, device=dev)
        frame.AssignFeaturesToGrid(self.d_kps, self.d_counts, self.bounds, self.d_cs, self.d_ci)
        self.frames = search.device_frames(self.d_kps, self.d_desc, self.d_counts, self.bounds, self.d_cs, self.d_ci, self.d_ur, self.d_occ)
        self.d_scratch = torch.zeros(search.scratch_bytes(B, qcap, cap), dtype=torch.uint8, device=dev)
        self.d_mk = torch.full((B, cap), -7, dtype=torch.int32, device=dev)
        self.d_mq = torch.full((B, qcap), -7, dtype=torch.int32, device=dev)
        self.d_nm = torch.full((B,), -7, dtype=torch.int32, device=dev)
        self.counts, self.qcounts = counts, qcounts


def scenes_ragged(seed0, rep=1, **kw):
    sc = [tracking_scene(2000, 1900, seed0, **kw), tracking_scene(1337, 2100, seed0 + 1, **kw), tracking_scene(700, 650, seed0 + 2, **kw),
          tracking_scene(900, 800, seed0 + 3, **kw), tracking_scene(800, 900, seed0 + 4, **kw)]
    # frame 3: no keypoints in the current frame; frame 4: no map points
    for k in ("cur", "cur_desc", "uright", "occupied0"):
        sc[3][k] = sc[3][k][:0]
    for k in ("last", "Xw", "mp_flags", "mp_desc"):
        sc[4][k] = sc[4][k][:0]
    return sc * rep   # rep = 2: ten pairs per call - the batch form of the vocabulary ordering (one CTA per pair)


@pytest.mark.parametrize("seed,th,direction", [(100, 15.0, 0), (200, 7.0, 1), (300, 15.0, 2), (400, 40.0, 0)])
def test_last_frame_search(oracle, seed, th, direction):
    import torch
    from orb_slam2_detailed_comments_b200 import search
    from orb_slam2_detailed_comments_b200._lib import PROJ_QUERY_DTYPE
    sc = scenes_ragged(seed, frac_unobserved=0.3 if seed == 400 else 0.05)
    bt = Batch(sc, 2100, 2200)
    B = len(sc)
    d_q = torch.zeros((B, bt.qcap, 32), dtype=torch.uint8, device="cuda")
    d_dir = torch.full((B,), direction, dtype=torch.int32, device="cuda")
    search.ProjectLastFrame(bt.d_Xw, bt.d_fl, bt.d_last, bt.d_qcounts, bt.d_T, d_dir, sc[0]["cam4"], bt.bounds, sc[0]["mbf"], th, SF, d_q)
    search.SearchByProjection(bt.frames, d_q, bt.d_mpd, bt.d_qcounts, search.ORB_SEARCH_BEST, search.TH_HIGH, 0.9, True, bt.d_scratch,
                              bt.d_mk, bt.d_mq, bt.d_nm)
    torch.cuda.synchronize()
    q = d_q.cpu().numpy().view(PROJ_QUERY_DTYPE).reshape(B, bt.qcap)
    mk, mq, nm = bt.d_mk.cpu().numpy(), bt.d_mq.cpu().numpy(), bt.d_nm.cpu().numpy()
    total = 0
    for b, s in enumerate(sc):
        n, m = bt.counts[b], bt.qcounts[b]
        qr = oracle.project_last_frame(s["Xw"], s["mp_flags"], s["last"], s["Tcw"], s["cam4"], s["bounds"], s["mbf"], th, SF, direction)
        assert q[b, :m].tobytes() == qr.tobytes(), "projected queries differ in frame %d" % b
        rn, rk, rq = oracle.search_by_projection(s["cur"], s["cur_desc"], s["uright"], s["bounds"], s["occupied0"], qr, s["mp_desc"],
                                                 oracle.SEARCH_BEST, 100, 0.9, True)
        assert nm[b] == rn, "nmatches differ in frame %d: %d vs %d" % (b, nm[b], rn)
        assert np.array_equal(mk[b, :n], rk) and np.all(mk[b, n:] == -1)
        assert np.array_equal(mq[b, :m], rq) and np.all(mq[b, m:] == -1)
        total += rn
    assert total > 1500


@pytest.mark.parametrize("seed,th", [(500, 1.0), (600, 3.0), (700, 5.0)])
def test_local_map_search(oracle, seed, th):
    import torch
    from orb_slam2_detailed_comments_b200 import search
    sc = scenes_ragged(seed, frac_mapped=0.9)
    bt = Batch(sc, 2100, 2200)
    B = len(sc)
    rng = np.random.RandomState(seed)
    qs = np.zeros((B, bt.qcap), search.PROJ_QUERY_DTYPE)
    for b, s in enumerate(sc):
        m = bt.qcounts[b]
        q0 = oracle.project_last_frame(s["Xw"], s["mp_flags"] | 1, s["last"], s["Tcw"], s["cam4"], s["bounds"], s["mbf"], 1.0, SF, 0)
        in_view = (q0["flags"] & 1) & (rng.rand(m) > 0.03)
        qs[b, :m] = search.local_map_queries(q0["u"], q0["v"], q0["ur"], s["last"]["octave"], 0.99 + 0.01 * rng.rand(m), in_view,
                                             (s["mp_flags"] >> 1) & 1, th, SF)
    d_q = torch.from_numpy(qs.view(np.uint8).reshape(B, bt.qcap, 32)).cuda()
    search.SearchByProjection(bt.frames, d_q, bt.d_mpd, bt.d_qcounts, search.ORB_SEARCH_RATIO_LEVEL, search.TH_HIGH, 0.8, False,
                              bt.d_scratch, bt.d_mk, bt.d_mq, bt.d_nm)
    torch.cuda.synchronize()
    mk, mq, nm = bt.d_mk.cpu().numpy(), bt.d_mq.cpu().numpy(), bt.d_nm.cpu().numpy()
    total = 0
    for b, s in enumerate(sc):
        n, m = bt.counts[b], bt.qcounts[b]
        rn, rk, rq = oracle.search_by_projection(s["cur"], s["cur_desc"], s["uright"], s["bounds"], s["occupied0"], qs[b, :m], s["mp_desc"],
                                                 oracle.SEARCH_RATIO_LEVEL, 100, 0.8, False)
        assert nm[b] == rn
        assert np.array_equal(mk[b, :n], rk) and np.array_equal(mq[b, :m], rq)
        total += rn
    assert total > 1000


@pytest.mark.parametrize("seed,nodes,ori,rep", [(800, 100, True, 1), (900, 12, True, 1), (1000, 3000, False, 1), (810, 100, True, 2)])
def test_search_by_bow(oracle, seed, nodes, ori, rep):
    import torch
    from orb_slam2_detailed_comments_b200 import search
    sc = scenes_ragged(seed, rep, flip_bits=40)
    bt = Batch(sc, 2100, 2200)
    B = len(sc)
    rng = np.random.RandomState(seed)
    node1 = np.full((B, bt.qcap), -1, np.int32); node2 = np.full((B, bt.cap), -1, np.int32)
    for b, s in enumerate(sc):
        n, m = bt.counts[b], bt.qcounts[b]
        node2[b, :n] = rng.randint(0, nodes, n) * 37 + 5
        if n and m:
            node1[b, :m] = np.where(rng.rand(m) < 0.85, node2[b, :n][s["src"][:m] % n], rng.randint(0, nodes, m) * 37 + 5)
        node1[b, :m][rng.rand(m) < 0.02] = -1
        node2[b, :n][rng.rand(n) < 0.02] = -1
    d_n1 = torch.from_numpy(node1).cuda(); d_n2 = torch.from_numpy(node2).cuda()
    d_us = (bt.d_fl & 1).contiguous()
    search.SearchByBoW(bt.d_last, bt.d_mpd, d_n1, d_us, bt.d_qcounts, bt.frames, d_n2, 0.7, ori, bt.d_scratch, bt.d_mk, bt.d_mq, bt.d_nm)
    torch.cuda.synchronize()
    mk, mq, nm = bt.d_mk.cpu().numpy(), bt.d_mq.cpu().numpy(), bt.d_nm.cpu().numpy()
    total = 0
    for b, s in enumerate(sc):
        n, m = bt.counts[b], bt.qcounts[b]
        rn, rk, rq = oracle.search_by_bow(s["last"], s["mp_desc"], node1[b, :m], s["mp_flags"] & 1, s["cur"], s["cur_desc"], node2[b, :n],
                                          50, 0.7, ori, unusable2=s["occupied0"])   # bt.frames carries the occupancy
        assert nm[b] == rn
        assert np.array_equal(mk[b, :n], rk) and np.array_equal(mq[b, :m], rq)
        total += rn
    assert total > 300


def test_search_by_bow_keyframe_pair(oracle):
    """SearchByBoW(KeyFrame*, KeyFrame*) (ORBmatcher.cc:729) through the same entry point: unusable keyframe-2 features
    as initial occupancy, th = TH_LOW - 1."""
    import torch
    from orb_slam2_detailed_comments_b200 import search
    sc = scenes_ragged(1100, flip_bits=60)
    bt = Batch(sc, 2100, 2200)
    B = len(sc)
    rng = np.random.RandomState(11)
    node1 = np.full((B, bt.qcap), -1, np.int32); node2 = np.full((B, bt.cap), -1, np.int32)
    unusable2 = np.zeros((B, bt.cap), np.uint8)
    for b, s in enumerate(sc):
        n, m = bt.counts[b], bt.qcounts[b]
        node2[b, :n] = rng.randint(0, 60, n)
        if n and m:
            node1[b, :m] = np.where(rng.rand(m) < 0.85, node2[b, :n][s["src"][:m] % n], rng.randint(0, 60, m))
        unusable2[b, :n] = rng.rand(n) < 0.3
    d_n1 = torch.from_numpy(node1).cuda(); d_n2 = torch.from_numpy(node2).cuda()
    d_us = (bt.d_fl & 1).contiguous()
    d_un2 = torch.from_numpy(unusable2).cuda()
    frames = search.device_frames(bt.d_kps, bt.d_desc, bt.d_counts, bt.bounds, None, None, None, d_un2)
    search.SearchByBoW(bt.d_last, bt.d_mpd, d_n1, d_us, bt.d_qcounts, frames, d_n2, 0.8, True, bt.d_scratch, bt.d_mk, bt.d_mq, bt.d_nm,
                       th=search.TH_LOW - 1)
    torch.cuda.synchronize()
    mk, mq, nm = bt.d_mk.cpu().numpy(), bt.d_mq.cpu().numpy(), bt.d_nm.cpu().numpy()
    total = 0
    for b, s in enumerate(sc):
        n, m = bt.counts[b], bt.qcounts[b]
        rn, rk, rq = oracle.search_by_bow(s["last"], s["mp_desc"], node1[b, :m], s["mp_flags"] & 1, s["cur"], s["cur_desc"], node2[b, :n],
                                          49, 0.8, True, unusable2=unusable2[b, :n])
        assert nm[b] == rn and np.array_equal(mk[b, :n], rk) and np.array_equal(mq[b, :m], rq)
        total += rn
    assert total > 200


@pytest.mark.parametrize("seed,only_stereo,mono,rep", [(1200, 0, False, 1), (1300, 1, False, 1), (1400, 0, True, 1), (1210, 0, False, 2)])
def test_search_for_triangulation(oracle, seed, only_stereo, mono, rep):
    import torch
    from orb_slam2_detailed_comments_b200 import search
    from orb_slam2_detailed_comments_b200._lib import TRI_PAIR_DTYPE
    from orb_slam2_detailed_comments_b200.synth import triangulation_pair
    sc = [dict(s) for s in scenes_ragged(seed, rep, flip_bits=50, noise_px=1.0)]
    tps = [triangulation_pair(s, seed + i) for i, s in enumerate(sc)]
    for s, tp in zip(sc, tps):          # keyframe 1 = the scene's map-point side, seen from the identity pose
        tp["kps1"] = tp["kps1"][:len(s["last"])]; tp["has_mp1"] = tp["has_mp1"][:len(s["last"])]; tp["ur1"] = tp["ur1"][:len(s["last"])]
        tp["has_mp2"] = tp["has_mp2"][:len(s["cur"])]
        s["last"] = tp["kps1"]
    bt = Batch(sc, 2100, 2200)
    B = len(sc)
    rng = np.random.RandomState(seed)
    node1 = np.full((B, bt.qcap), -1, np.int32); node2 = np.full((B, bt.cap), -1, np.int32)
    has1 = np.zeros((B, bt.qcap), np.uint8); has2 = np.zeros((B, bt.cap), np.uint8); ur1 = np.full((B, bt.qcap), -1, np.float32)
    pairs = np.zeros(B, TRI_PAIR_DTYPE)
    for b, (s, tp) in enumerate(zip(sc, tps)):
        n, m = bt.counts[b], bt.qcounts[b]
        node2[b, :n] = rng.randint(0, 80, n)
        if n and m:
            node1[b, :m] = np.where(rng.rand(m) < 0.85, node2[b, :n][s["src"][:m] % n], rng.randint(0, 80, m))
        has1[b, :m], has2[b, :n], ur1[b, :m] = tp["has_mp1"], tp["has_mp2"], tp["ur1"]
        pairs["F12"][b] = tp["F12"]; pairs["ex"][b], pairs["ey"][b], pairs["only_stereo"][b] = tp["ex"], tp["ey"], only_stereo
    SIG = (SF * SF).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_has2 = t(has2)
    frames2 = search.device_frames(bt.d_kps, bt.d_desc, bt.d_counts, bt.bounds, None, None, None if mono else bt.d_ur, d_has2)
    d_m12 = torch.full((B, bt.qcap), -7, dtype=torch.int32, device="cuda")
    search.SearchForTriangulation(bt.d_last, bt.d_mpd, t(node1), t(has1), None if mono else t(ur1), bt.d_qcounts, frames2, t(node2),
                                  t(pairs.view(np.uint8).reshape(B, 48)), SF, SIG, True, bt.d_scratch, d_m12, bt.d_nm)
    torch.cuda.synchronize()
    m12, nm = d_m12.cpu().numpy(), bt.d_nm.cpu().numpy()
    total = 0
    for b, (s, tp) in enumerate(zip(sc, tps)):
        n, m = bt.counts[b], bt.qcounts[b]
        rn, r12 = oracle.search_for_triangulation(s["last"], s["mp_desc"], node1[b, :m], has1[b, :m], None if mono else ur1[b, :m], s["cur"],
                                                  s["cur_desc"], node2[b, :n], has2[b, :n], None if mono else s["uright"], pairs[b], SF, SIG, True)
        assert nm[b] == rn, (b, nm[b], rn)
        assert np.array_equal(m12[b, :m], r12) and np.all(m12[b, m:] == -1)
        total += rn
    assert total > (60 if only_stereo else 200), total


@pytest.mark.parametrize("seed,th,direction", [(110, 15.0, 0), (210, 7.0, 1)])
def test_last_frame_search_against_the_reference_itself(reference, seed, th, direction):
    """The tracking matcher on the GPU (projection + SearchByProjection(CurrentFrame, LastFrame)) against the reference's own
    ORBmatcher.cc running on live ORB_SLAM2::Frame / MapPoint objects (oracle/_ref/liborbref.so, compiled unmodified):
    CurrentFrame.mvpMapPoints identical, frame by frame of a ragged batch."""
    import torch
    from orb_slam2_detailed_comments_b200 import search
    sc = scenes_ragged(seed)
    bt = Batch(sc, 2100, 2200)
    B = len(sc)
    d_q = torch.zeros((B, bt.qcap, 32), dtype=torch.uint8, device="cuda")
    d_dir = torch.full((B,), direction, dtype=torch.int32, device="cuda")
    search.ProjectLastFrame(bt.d_Xw, bt.d_fl, bt.d_last, bt.d_qcounts, bt.d_T, d_dir, sc[0]["cam4"], bt.bounds, sc[0]["mbf"], th, SF, d_q)
    search.SearchByProjection(bt.frames, d_q, bt.d_mpd, bt.d_qcounts, search.ORB_SEARCH_BEST, search.TH_HIGH, 0.9, True, bt.d_scratch,
                              bt.d_mk, bt.d_mq, bt.d_nm)
    torch.cuda.synchronize()
    mk, nm = bt.d_mk.cpu().numpy(), bt.d_nm.cpu().numpy()
    total = 0
    for b, s in enumerate(sc):
        n = bt.counts[b]
        if n == 0:
            assert nm[b] == 0
            continue
        cam9 = np.concatenate([s["cam4"], np.zeros(5, np.float32)])
        F = reference.ReferenceFrame(s["cur"], s["cur_desc"], cam9, 1241, 376)
        rn, rmk = reference.search_last_frame(F, s["uright"], s["occupied0"], s["last"], s["Xw"], s["mp_flags"], s["mp_desc"], s["Tcw"],
                                              s["cam4"], s["mbf"], s["mb"], th, direction, SF)
        assert nm[b] == rn, "nmatches differ in frame %d: %d vs %d" % (b, nm[b], rn)
        assert np.array_equal(mk[b, :n], rmk)
        total += rn
    print("seed", seed, "matches", total)
    assert total > 1500
