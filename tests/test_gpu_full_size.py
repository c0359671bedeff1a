"""Parity at BASELINE.json's full sizes through size-independent properties: the 1024-stereo-pair
KITTI batch and the 4096-pair matching batch are built from small pools, so (i) every repetition of
a pool element must give byte-identical output wherever it sits in the batch / chunk / lane, and
(ii) the first pool elements are checked against the oracle."""
import zlib

import numpy as np
import pytest

from conftest import CONFIGS

pytestmark = pytest.mark.gpu


def test_kitti_batch_of_1024_pairs(oracle):
    import torch
    from orb_slam2_detailed_comments_b200 import KP_DTYPE, ORBextractor
    from orb_slam2_detailed_comments_b200.synth import synth_batch
    w, h, nfeat = CONFIGS["kitti"]
    pool_n, total = 24, 2048
    pool = synth_batch(w, h, pool_n, seed0=900)
    gpu = ORBextractor(nfeat, 1.2, 8, 20, 7, max_batch=256)
    cap = gpu.max_keypoints
    idx = np.arange(total) % pool_n
    d_imgs = torch.from_numpy(pool).cuda()[torch.from_numpy(idx).cuda()].contiguous()
    d_kps = torch.zeros((total, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((total, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros(total, dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    gpu.extract_batch_device(d_imgs, d_kps, d_desc, d_counts, stream=ts.cuda_stream)
    gpu.synchronize(ts.cuda_stream)
    counts = d_counts.cpu().numpy()
    kps = d_kps.cpu().numpy(); desc = d_desc.cpu().numpy()
    # (i) checksum per frame over the valid records; all repetitions of a pool frame agree
    crc = np.array([zlib.crc32(kps[i, :counts[i]].tobytes() + desc[i, :counts[i]].tobytes()) for i in range(total)], np.uint32)
    for p in range(pool_n):
        assert len(set(crc[p::pool_n].tolist())) == 1, "frame %d differs between repetitions" % p
        assert len(set(counts[p::pool_n].tolist())) == 1
    # (ii) the pool itself against the oracle
    orc = oracle.OracleExtractor(nfeat, 1.2, 8, 20, 7)
    for p in range(4):
        okps, odesc = orc(pool[p])
        n = counts[p]
        k = kps[p, :n].view(KP_DTYPE).reshape(-1)
        assert n == len(okps)
        for f in ("x", "y", "octave", "response", "size"):
            assert np.array_equal(k[f], okps[f])
        assert np.abs(k["angle"] - okps["angle"]).max() <= 1e-3
        assert (desc[p, :n] == odesc).all(1).mean() >= 0.999
    assert counts.min() >= nfeat


def test_matching_batch_of_4096_pairs(oracle):
    import torch
    from orb_slam2_detailed_comments_b200 import ORBmatcher
    from orb_slam2_detailed_comments_b200.synth import correlated_descriptor_pair
    pool_n, total, n = 16, 4096, 2000
    desc = np.zeros((2 * pool_n, n, 32), np.uint8); ang = np.zeros((2 * pool_n, n), np.float32)
    for p in range(pool_n):
        A, B, aa, ab = correlated_descriptor_pair(n, 4000 + p)
        desc[2 * p], desc[2 * p + 1], ang[2 * p], ang[2 * p + 1] = A, B, aa, ab
    reps = total // pool_n
    d_desc = torch.from_numpy(desc).cuda().repeat(reps, 1, 1).contiguous()
    d_ang = torch.from_numpy(ang).cuda().repeat(reps, 1).contiguous()
    d_m12 = torch.zeros((total, n), dtype=torch.int32, device="cuda")
    d_nm = torch.zeros(total, dtype=torch.int32, device="cuda")
    m = ORBmatcher(0.9, True)
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    m.match_pairs_device(d_desc, d_ang, d_m12, d_nm, stream=ts.cuda_stream)
    m.synchronize(ts.cuda_stream)
    m12 = d_m12.cpu().numpy(); nm = d_nm.cpu().numpy()
    for p in range(pool_n):
        assert (m12[p::pool_n] == m12[p]).all() and (nm[p::pool_n] == nm[p]).all()
    xy = np.zeros((n, 2), np.float32); oc = np.zeros(n, np.int32)
    for p in range(3):
        n_ref, m_ref, _, _, _ = oracle.search_for_initialization(xy, oc, ang[2 * p], desc[2 * p], xy, oc, ang[2 * p + 1],
                                                                 desc[2 * p + 1], (0, 1, 0, 1), xy, nnratio=0.9,
                                                                 check_ori=True, mode=1)
        assert nm[p] == n_ref and np.array_equal(m12[p], m_ref)
    # every match index is a valid column and no column is used twice (the dedup invariant)
    for p in range(pool_n):
        used = m12[p][m12[p] >= 0]
        assert used.max() < n and len(np.unique(used)) == len(used)
