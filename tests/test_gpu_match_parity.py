"""GPU parity of the matcher kernels (through the C ABI) against the CPU oracle: Hamming
distances, 2-NN best/second, match indices and counts are bit-exact integers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames_from_extraction(oracle, shift=(6, 3), seed=21, nfeat=1000):
    """Two real descriptor sets: a synthetic frame and a shifted, re-noised copy."""
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    big = synth_frame(700, 520, seed, noise_sigma=0.0).astype(np.float32)
    rng = np.random.RandomState(seed)
    a = np.clip(np.rint(big[20:500, 20:660] + rng.normal(0, 2.0, (480, 640))), 0, 255).astype(np.uint8)
    b = np.clip(np.rint(big[20 + shift[1]:500 + shift[1], 20 + shift[0]:660 + shift[0]] + rng.normal(0, 2.0, (480, 640))), 0, 255).astype(np.uint8)
    orc = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    ka, da = orc(a)
    kb, db = orc(b)
    return ka, da, kb, db


def test_hamming_matrix(oracle):
    import torch
    from orb_slam2_detailed_comments_b200 import ORBmatcher
    from orb_slam2_detailed_comments_b200.synth import random_descriptors
    a = random_descriptors(300, 1); b = random_descriptors(257, 2)
    b[:10] = a[:10]          # zero distances
    b[10] = ~a[10]           # distance 256
    m = ORBmatcher(0.9, True)
    d_out = torch.zeros((300, 257), dtype=torch.int32, device="cuda")
    m.hamming_matrix_device(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), d_out)
    m.synchronize()
    ref = oracle.hamming_matrix(a, b)
    assert np.array_equal(d_out.cpu().numpy(), ref)
    assert ref[10, 10] == 256 and ref[0, 0] == 0
    assert ORBmatcher.DescriptorDistance(a[3], b[200]) == ref[3, 200] == oracle.hamming(a[3], b[200])


@pytest.mark.parametrize("window", [100, 30])
def test_search_for_initialization_reference_mode(oracle, window):
    from orb_slam2_detailed_comments_b200 import FrameView, ORBmatcher
    ka, da, kb, db = _frames_from_extraction(oracle)
    F1 = FrameView.from_keypoints(ka, da, 640, 480)
    F2 = FrameView.from_keypoints(kb, db, 640, 480)
    prev = F1.xy.copy()
    n_ref, m_ref, prev_ref, best_ref, second_ref = oracle.search_for_initialization(
        F1.xy, F1.octave, F1.angle, F1.descriptors, F2.xy, F2.octave, F2.angle, F2.descriptors,
        (0, 640, 0, 480), prev, window=window, nnratio=0.9, check_ori=True, mode=0)
    m = ORBmatcher(0.9, True)
    prev_gpu = prev.copy()
    n, m12, best, second = m.SearchForInitialization(F1, F2, prev_gpu, window, mode=0, want_distances=True)
    print("window", window, "matches", n_ref, "rows", F1.N)
    assert n_ref > 20
    assert np.array_equal(best, best_ref) and np.array_equal(second, second_ref)
    assert np.array_equal(m12, m_ref) and n == n_ref
    assert np.array_equal(prev_gpu, prev_ref)
    # second call with the updated vbPrevMatched, as Tracking does on the next frame
    n2_ref, m2_ref, prev2_ref, _, _ = oracle.search_for_initialization(
        F1.xy, F1.octave, F1.angle, F1.descriptors, F2.xy, F2.octave, F2.angle, F2.descriptors,
        (0, 640, 0, 480), prev_ref, window=window, nnratio=0.9, check_ori=True, mode=0)
    n2, m2 = m.SearchForInitialization(F1, F2, prev_gpu, window, mode=0)
    assert n2 == n2_ref and np.array_equal(m2, m2_ref) and np.array_equal(prev_gpu, prev2_ref)


def _clustered_level0_frames(n, seed, w=640, h=480, clusters=6, flips=6):
    """Two frames of level-0 keypoints whose descriptors come from a few prototypes (repetitive texture): most rows see many
    candidates at small distances, rows steal each other's matches, candidate lists overflow."""
    from orb_slam2_detailed_comments_b200._lib import KP_DTYPE
    rng = np.random.RandomState(seed)
    proto = rng.randint(0, 256, (clusters, 32)).astype(np.uint8)

    def frame(m, jitter):
        k = np.zeros(m, KP_DTYPE)
        k["x"] = rng.uniform(20, w - 20, m).astype(np.float32); k["y"] = rng.uniform(20, h - 20, m).astype(np.float32)
        k["octave"] = (rng.rand(m) < 0.15).astype(np.int32) * rng.randint(1, 4, m)    # 85 % on level 0
        k["angle"] = rng.uniform(0, 360, m).astype(np.float32)
        k["size"] = 31.0; k["class_id"] = -1
        d = proto[rng.randint(0, clusters, m)].copy()
        for i in range(m):
            bits = rng.randint(0, 256, rng.randint(0, flips + 1))
            np.bitwise_xor.at(d[i], bits >> 3, (1 << (bits & 7)).astype(np.uint8))
        return k, d
    ka, da = frame(n, 0)
    kb, db = frame(n + 37, 1)
    m = min(n, 200)
    kb["x"][:m] = ka["x"][:m] + 3.0; kb["y"][:m] = ka["y"][:m] - 2.0; kb["octave"][:m] = ka["octave"][:m]
    db[:m] = da[:m]                                     # exact duplicates: distance 0, ties between rows
    return ka, da, kb, db


@pytest.mark.parametrize("case", ["extraction", "extraction_narrow", "clustered", "clustered_wide", "ratio_1", "ratio_small", "tiny"])
def test_search_for_initialization_latency_form(oracle, case):
    """The form a drop-in caller gets (no per-row distances requested): k_sfi_scan + k_sfi_walk. Same vnMatches12, return
    value and vbPrevMatched as the oracle, over three consecutive calls (vbPrevMatched moves, Tracking.cc:915-926), on real
    extractions and on repetitive descriptors that overflow the stored candidate lists (exact re-scoring path)."""
    from orb_slam2_detailed_comments_b200 import FrameView, ORBmatcher
    ratio, window, ori = 0.9, 100, True
    if case.startswith("extraction"):
        ka, da, kb, db = _frames_from_extraction(oracle, nfeat=2000 if case == "extraction" else 1000)
        window = 100 if case == "extraction" else 12
    elif case == "tiny":
        ka, da, kb, db = _clustered_level0_frames(9, 3)
    else:
        ka, da, kb, db = _clustered_level0_frames(1500 if case != "clustered_wide" else 900, 11, clusters=4 if case == "clustered" else 2,
                                                  flips=10 if case == "clustered" else 4)
        window = 100 if case == "clustered" else 400
        ratio = {"ratio_1": 1.0, "ratio_small": 0.3}.get(case, 0.9)
        ori = case != "ratio_1"
    F1 = FrameView.from_keypoints(ka, da, 640, 480)
    F2 = FrameView.from_keypoints(kb, db, 640, 480)
    m = ORBmatcher(ratio, ori)
    prev_ref = F1.xy.copy()
    prev_gpu = F1.xy.copy()
    total = 0
    for call in range(3):
        n_ref, m_ref, prev_ref, _, _ = oracle.search_for_initialization(
            F1.xy, F1.octave, F1.angle, F1.descriptors, F2.xy, F2.octave, F2.angle, F2.descriptors, (0, 640, 0, 480), prev_ref,
            window=window, nnratio=ratio, check_ori=ori, mode=0)
        n, m12 = m.SearchForInitialization(F1, F2, prev_gpu, window, mode=0)
        assert n == n_ref, "%s call %d: %d matches vs %d" % (case, call, n, n_ref)
        assert np.array_equal(m12, m_ref), "%s call %d: vnMatches12 differs in %d rows" % (case, call, int((m12 != m_ref).sum()))
        assert np.array_equal(prev_gpu, prev_ref)
        total += n_ref
    print(case, "matches over three calls", total)
    if case != "tiny":
        assert total > 30
    m.close()


@pytest.mark.parametrize("n", [2600, 2048, 2000, 1000, 777, 130, 33])
@pytest.mark.parametrize("check_ori", [True, False])
def test_bruteforce_single_pair(oracle, n, check_ori):
    from orb_slam2_detailed_comments_b200 import FrameView, ORBmatcher
    from orb_slam2_detailed_comments_b200.synth import correlated_descriptor_pair
    A, B, angA, angB = correlated_descriptor_pair(n, 5 + n)
    xy = np.zeros((n, 2), np.float32); oc = np.zeros(n, np.int32)
    F1 = FrameView(xy, oc, angA, A, (0, 1, 0, 1)); F2 = FrameView(xy, oc, angB, B, (0, 1, 0, 1))
    n_ref, m_ref, _, best_ref, second_ref = oracle.search_for_initialization(
        xy, oc, angA, A, xy, oc, angB, B, (0, 1, 0, 1), xy, nnratio=0.9, check_ori=check_ori, mode=1)
    m = ORBmatcher(0.9, check_ori)
    nm, m12, best, second = m.SearchForInitialization(F1, F2, xy.copy(), 0, mode=1, want_distances=True)
    print("n", n, "matches", n_ref)
    assert n_ref > n // 4
    assert np.array_equal(best, best_ref) and np.array_equal(second, second_ref)
    assert nm == n_ref and np.array_equal(m12, m_ref)


def test_dedup_steal_path(oracle):
    """Many rows compete for few columns: exercises the vMatchedDistance skip and the steal
    (ORBmatcher.cc:627, :650-654)."""
    from orb_slam2_detailed_comments_b200 import FrameView, ORBmatcher
    rng = np.random.RandomState(3)
    base = rng.randint(0, 256, (40, 32)).astype(np.uint8)
    bits = np.unpackbits(base, axis=1)
    rows = []
    for r in range(600):
        src = bits[r % 40].copy()
        flip = rng.rand(256) < (0.01 + 0.1 * rng.rand())
        rows.append(np.packbits(src ^ flip.astype(np.uint8)))
    A = np.stack(rows)
    B = np.concatenate([base, rng.randint(0, 256, (300, 32)).astype(np.uint8)])
    n1, n2 = len(A), len(B)
    a1 = (rng.rand(n1) * 360).astype(np.float32); a2 = (rng.rand(n2) * 360).astype(np.float32)
    z1 = np.zeros((n1, 2), np.float32); z2 = np.zeros((n2, 2), np.float32)
    o1 = np.zeros(n1, np.int32); o2 = np.zeros(n2, np.int32)
    n_ref, m_ref, _, b_ref, s_ref = oracle.search_for_initialization(z1, o1, a1, A, z2, o2, a2, B, (0, 1, 0, 1), z1,
                                                                     nnratio=0.9, check_ori=False, mode=1)
    m = ORBmatcher(0.9, False)
    nm, m12, b, s = m.SearchForInitialization(FrameView(z1, o1, a1, A, (0, 1, 0, 1)), FrameView(z2, o2, a2, B, (0, 1, 0, 1)),
                                              z1.copy(), 0, mode=1, want_distances=True)
    assert np.array_equal(b, b_ref) and np.array_equal(s, s_ref)
    assert nm == n_ref and np.array_equal(m12, m_ref)
    assert (m_ref >= 0).sum() <= 40 and n_ref > 10


def test_pairs_batch_device(oracle):
    import torch
    from orb_slam2_detailed_comments_b200 import ORBmatcher
    from orb_slam2_detailed_comments_b200.synth import correlated_descriptor_pair
    P, n = 12, 2000
    desc = np.zeros((2 * P, n, 32), np.uint8); ang = np.zeros((2 * P, n), np.float32)
    for p in range(P):
        A, B, aa, ab = correlated_descriptor_pair(n, 1000 + p)
        desc[2 * p], desc[2 * p + 1], ang[2 * p], ang[2 * p + 1] = A, B, aa, ab
    m = ORBmatcher(0.9, True)
    d_m12 = torch.zeros((P, n), dtype=torch.int32, device="cuda")
    d_nm = torch.zeros(P, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    m.match_pairs_device(torch.from_numpy(desc).cuda(), torch.from_numpy(ang).cuda(), d_m12, d_nm, stream=s)
    m.synchronize(s)
    m12 = d_m12.cpu().numpy(); nm = d_nm.cpu().numpy()
    xy = np.zeros((n, 2), np.float32); oc = np.zeros(n, np.int32)
    for p in range(P):
        n_ref, m_ref, _, _, _ = oracle.search_for_initialization(xy, oc, ang[2 * p], desc[2 * p], xy, oc, ang[2 * p + 1],
                                                                 desc[2 * p + 1], (0, 1, 0, 1), xy, nnratio=0.9,
                                                                 check_ori=True, mode=1)
        assert nm[p] == n_ref and np.array_equal(m12[p], m_ref)
    total, nm_mt = oracle.match_batch_mt(desc, ang, 0.9, nthreads=2)
    assert np.array_equal(nm_mt, nm)


def test_allpairs_counts(oracle):
    import torch
    from orb_slam2_detailed_comments_b200 import ORBmatcher
    from orb_slam2_detailed_comments_b200.synth import correlated_descriptor_pair
    nkf, nd = 10, 300
    all_desc = np.zeros((nkf, nd, 32), np.uint8)
    A, _, _, _ = correlated_descriptor_pair(nd, 77)
    rng = np.random.RandomState(9)
    bits = np.unpackbits(A, axis=1)
    for k in range(nkf):
        flips = (rng.rand(*bits.shape) < 0.03 * (k % 4)).astype(np.uint8)
        all_desc[k] = np.packbits(bits ^ flips, axis=1)[rng.permutation(nd)]
    ref = oracle.allpairs_counts(all_desc, 0.9)
    m = ORBmatcher(0.9, True)
    d_all = torch.from_numpy(all_desc).cuda()
    d_counts = torch.zeros((nkf, nkf), dtype=torch.int32, device="cuda")
    m.match_allpairs_device(d_all, 0, nkf, d_counts)
    d_cols = torch.zeros((nkf, nkf), dtype=torch.int32, device="cuda")
    m.match_allpairs_device(d_all, 0, nkf, d_cols, col_begin=0, col_end=4)
    m.match_allpairs_device(d_all, 0, nkf, d_cols, col_begin=4, col_end=nkf)
    m.synchronize()
    assert np.array_equal(d_counts.cpu().numpy(), ref)
    assert np.array_equal(d_cols.cpu().numpy(), ref)     # same matrix assembled from two column blocks
    # a row block, as one rank of the sharded workload computes it
    d_part = torch.zeros((3, nkf), dtype=torch.int32, device="cuda")
    m.match_allpairs_device(d_all, 4, 7, d_part)
    m.synchronize()
    assert np.array_equal(d_part.cpu().numpy(), ref[4:7])
    assert ref.max() > 50


def test_int_pipe_peak_runs():
    from orb_slam2_detailed_comments_b200 import int_pipe_peak
    pk = int_pipe_peak(0)
    print(pk)
    assert pk["popc"] > 1e11 and pk["lop3"] > pk["popc"]


@pytest.mark.parametrize("window", [100, 30])
def test_search_for_initialization_against_the_reference_itself(oracle, reference, window):
    """The CUDA matcher against the reference's own ORBmatcher::SearchForInitialization + Frame::GetFeaturesInArea
    (src/ORBmatcher.cc and src/Frame.cc compiled unmodified into oracle/_ref/liborbref.so): vnMatches12, the match
    count and the updated vbPrevMatched are identical, also on the second call with the updated positions."""
    from orb_slam2_detailed_comments_b200 import FrameView, ORBmatcher
    ka, da, kb, db = _frames_from_extraction(oracle)
    cam = np.array([500, 500, 320, 240, 0, 0, 0, 0, 0], np.float32)   # no distortion: mvKeysUn = mvKeys, bounds = image
    R1 = reference.ReferenceFrame(ka, da, cam, 640, 480); R2 = reference.ReferenceFrame(kb, db, cam, 640, 480)
    F1 = FrameView.from_keypoints(ka, da, 640, 480); F2 = FrameView.from_keypoints(kb, db, 640, 480)
    m = ORBmatcher(0.9, True)
    prev_gpu = F1.xy.copy(); prev_ref = F1.xy.copy()
    for call in range(2):
        n_ref, m_ref, prev_ref = reference.search_for_initialization(R1, R2, prev_ref, window, 0.9, True)
        n, m12 = m.SearchForInitialization(F1, F2, prev_gpu, window, mode=0)
        print("window", window, "call", call, "matches", n_ref)
        assert n == n_ref > 20 and np.array_equal(m12, m_ref) and np.array_equal(prev_gpu, prev_ref)


def test_allpairs_library_entry_single_rank(oracle):
    """orb_match_allpairs_nccl with world = 1 (no communicator): the multi-GPU entry point degenerates to the local kernel."""
    import torch
    from orb_slam2_detailed_comments_b200 import ORBmatcher
    from orb_slam2_detailed_comments_b200.distributed import allpairs_match_counts_nccl
    rng = np.random.RandomState(5)
    nkf, nd = 7, 300
    base = rng.randint(0, 256, (nd, 32)).astype(np.uint8)
    bits = np.unpackbits(base, axis=1)
    all_np = np.stack([np.packbits(bits ^ (rng.rand(*bits.shape) < 0.02 * (k % 4)).astype(np.uint8), axis=1)[rng.permutation(nd)]
                       for k in range(nkf)])
    m = ORBmatcher(0.9, True)
    counts = allpairs_match_counts_nccl(m, None, torch.from_numpy(all_np).cuda(), nkf)
    m.synchronize()
    assert np.array_equal(counts.cpu().numpy(), oracle.allpairs_counts(all_np, 0.9, 0, nkf))
