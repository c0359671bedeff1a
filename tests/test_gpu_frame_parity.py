"""GPU parity of the Frame helpers (UndistortKeyPoints, AssignFeaturesToGrid, GetFeaturesInArea)
through the C ABI against the oracle: bit-exact undistorted coordinates (double arithmetic in
OpenCV's operation order), identical grid and identical, identically ordered query results."""
import numpy as np
import pytest

from test_oracle_frame import NODIST, TUM1, random_kps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cam9,w,h", [(TUM1, 640, 480), (NODIST, 1241, 376)])
def test_undistort_grid_area(oracle, cam9, w, h):
    import torch
    from orb_slam2_detailed_comments_b200 import KP_DTYPE, frame
    B, cap = 3, 2100
    counts = np.array([2000, 1337, 0], np.int32)
    kps = np.zeros((B, cap), KP_DTYPE)
    for f in range(B):
        kps[f, :counts[f]] = random_kps(oracle, counts[f], w, h, 10 + f)
    cam = frame.camera(*[float(v) for v in cam9])
    bounds = frame.ComputeImageBounds(cam, w, h)
    assert np.array_equal(bounds, oracle.image_bounds(cam9, w, h))
    d_kps = torch.from_numpy(kps.view(np.uint8).reshape(B, cap, 28)).cuda()
    d_counts = torch.from_numpy(counts).cuda()
    d_un = torch.zeros_like(d_kps)
    d_start = torch.zeros((B, 64 * 48 + 1), dtype=torch.int32, device="cuda")
    d_items = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
    frame.UndistortKeyPoints(d_kps, d_counts, cam, d_un)
    frame.AssignFeaturesToGrid(d_un, d_counts, bounds, d_start, d_items)
    rng = np.random.RandomState(5)
    nq = 300
    q = frame.make_queries(rng.randint(0, 2, nq), rng.rand(nq) * (w + 60) - 30, rng.rand(nq) * (h + 60) - 30,
                           rng.rand(nq) * 120 + 1, rng.choice([-1, 0, 2], nq), rng.choice([-1, 0, 5], nq))
    d_q = torch.from_numpy(q.view(np.uint8).reshape(nq, 24)).cuda()
    d_out = torch.zeros((nq, cap), dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(nq, dtype=torch.int32, device="cuda")
    frame.GetFeaturesInArea(d_un, bounds, d_start, d_items, d_q, d_out, d_cnt)
    torch.cuda.synchronize()
    un = d_un.cpu().numpy().view(KP_DTYPE).reshape(B, cap)
    start = d_start.cpu().numpy(); items = d_items.cpu().numpy()
    out = d_out.cpu().numpy(); cnt = d_cnt.cpu().numpy()
    refs = []
    for f in range(B):
        ref = oracle.undistort_keypoints(kps[f, :counts[f]], cam9)
        refs.append(ref)
        assert un[f, :counts[f]].tobytes() == ref.tobytes(), "undistorted keypoints differ in frame %d" % f
        rs, ri = oracle.assign_grid(ref, bounds)
        assert np.array_equal(start[f], rs) and np.array_equal(items[f, :rs[-1]], ri)
    nonempty = 0
    for i in range(nq):
        f = int(q["frame"][i])
        ref = oracle.features_in_area(refs[f], bounds, q["x"][i], q["y"][i], q["r"][i], int(q["min_level"][i]), int(q["max_level"][i]))
        assert cnt[i] == len(ref) and out[i, :cnt[i]].tolist() == ref.tolist()
        nonempty += len(ref) > 0
    assert nonempty > nq // 3
