"""GPU parity of the input stage (cvtColor to gray, remap rectification) and of ComputeDistinctiveDescriptors
through the C ABI: bit-exact against the oracle (itself pinned bit-exactly against cv2 in test_oracle_input.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("code,channels,blue_first", [(0, 3, False), (1, 3, True), (2, 4, False), (3, 4, True)])
@pytest.mark.parametrize("w,h", [(1241, 376), (640, 480), (37, 5)])
def test_cvt_gray(oracle, code, channels, blue_first, w, h):
    import torch
    from orb_slam2_detailed_comments_b200 import input as I
    rng = np.random.RandomState(w + code)
    B = 3
    img = rng.randint(0, 256, (B, h, w, channels)).astype(np.uint8)
    d_src = torch.from_numpy(img).cuda()
    d_gray = torch.zeros((B, h, w), dtype=torch.uint8, device="cuda")
    I.cvtColorGray(d_src, code, d_gray)
    torch.cuda.synchronize()
    got = d_gray.cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], oracle.cvt_gray(img[b], blue_first))
    # padded rows on both sides (views into wider buffers): the per-pixel path
    d_wide = torch.zeros((B, h + 1, w + 3, channels), dtype=torch.uint8, device="cuda")
    d_wide[:, :h, :w] = d_src
    d_gwide = torch.full((B, h, w + 7), 99, dtype=torch.uint8, device="cuda")
    I.cvtColorGray(d_wide[:, :h, :w], code, d_gwide[:, :, :w])
    torch.cuda.synchronize()
    gw = d_gwide.cpu().numpy()
    assert np.array_equal(gw[:, :, :w], got) and np.all(gw[:, :, w:] == 99)


@pytest.mark.parametrize("w,h", [(752, 480), (333, 211)])
def test_remap(oracle, w, h):
    import torch
    from orb_slam2_detailed_comments_b200 import input as I
    from orb_slam2_detailed_comments_b200.synth import synth_frame
    cv2 = pytest.importorskip("cv2")
    from test_oracle_input import rectify_maps
    B = 6   # two frame groups of the kernel, the second one partial
    imgs = np.stack([synth_frame(w, h, 7 + b) for b in range(B)])
    rng = np.random.RandomState(1)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    maps = [rectify_maps(w, h, 3),
            ((xx * 1.1 - 20 + np.round(rng.rand(h, w) * 64) / 32).astype(np.float32), (yy * 1.1 - 15 + np.round(rng.rand(h, w) * 4) / 2).astype(np.float32)),
            ((rng.rand(h, w) * (w + 8) - 4).astype(np.float32), (rng.rand(h, w) * (h + 8) - 4).astype(np.float32))]
    d_src = torch.from_numpy(imgs).cuda()
    for mx, my in maps:
        d_mx, d_my = torch.from_numpy(np.ascontiguousarray(mx)).cuda(), torch.from_numpy(np.ascontiguousarray(my)).cuda()
        d_dst = torch.zeros((B, h, w), dtype=torch.uint8, device="cuda")
        I.remap(d_src, d_mx, d_my, d_dst)
        torch.cuda.synchronize()
        got = d_dst.cpu().numpy()
        for b in range(B):
            assert np.array_equal(got[b], oracle.remap_linear(imgs[b], mx, my))
            assert np.array_equal(got[b], cv2.remap(imgs[b], mx, my, cv2.INTER_LINEAR))
        d_wide = torch.full((B, h, w + 5), 77, dtype=torch.uint8, device="cuda")    # padded destination rows
        I.remap(d_src, d_mx, d_my, d_wide[:, :, :w])
        torch.cuda.synchronize()
        gw = d_wide.cpu().numpy()
        assert np.array_equal(gw[:, :, :w], got) and np.all(gw[:, :, w:] == 77)


def test_distinctive_descriptors(oracle):
    import torch
    from orb_slam2_detailed_comments_b200 import input as I
    rng = np.random.RandomState(9)
    counts = np.concatenate([[1, 2, 3, 4, 7, 16, 31, 32, 33, 64, 100, 0, 5], rng.randint(0, 40, 3000)])
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    desc = rng.randint(0, 256, (offsets[-1], 32)).astype(np.uint8)
    for p, c in enumerate(counts):
        if c > 2:   # clustered observations with duplicates: ties in the medians
            base = desc[offsets[p]].copy()
            flips = rng.randint(0, 256, (c, 40)); nfl = rng.randint(0, 40, c)
            for i in range(c):
                d = base.copy()
                bits = flips[i, :nfl[i]]
                np.bitwise_xor.at(d, bits >> 3, (1 << (bits & 7)).astype(np.uint8))
                desc[offsets[p] + i] = d
            desc[offsets[p] + c - 1] = desc[offsets[p]]
    ref = oracle.distinctive_descriptors(desc, offsets)
    d_desc = torch.from_numpy(desc).cuda(); d_off = torch.from_numpy(offsets).cuda()
    d_best = torch.full((len(counts),), -9, dtype=torch.int32, device="cuda")
    d_bd = torch.zeros((len(counts), 32), dtype=torch.uint8, device="cuda")
    I.ComputeDistinctiveDescriptors(d_desc, d_off, 100, d_best, d_bd)
    torch.cuda.synchronize()
    got = d_best.cpu().numpy(); bd = d_bd.cpu().numpy()
    assert np.array_equal(got, ref)
    sel = np.nonzero(ref >= 0)[0]
    assert np.array_equal(bd[sel], desc[offsets[sel] + ref[sel]])
    # bound too small: flagged, not overrun
    I.ComputeDistinctiveDescriptors(d_desc, d_off, 50, d_best, None)
    torch.cuda.synchronize()
    got = d_best.cpu().numpy()
    assert np.all(got[counts > 50] == -2) and np.array_equal(got[counts <= 50], ref[counts <= 50])
