"""Pins the oracle's Frame helpers: undistortion against cv2.undistortPoints (bit-exact), the image
bounds, and the grid / GetFeaturesInArea against a direct Python restatement of Frame.cc:399-423,
590-698."""
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

TUM1 = np.array([517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
NODIST = np.array([718.856, 718.856, 607.1928, 185.2157, 0, 0, 0, 0, 0], np.float32)


def random_kps(oracle, n, w, h, seed):
    rng = np.random.RandomState(seed)
    k = np.zeros(n, oracle.KP_DTYPE)
    k["x"] = (rng.rand(n) * w).astype(np.float32); k["y"] = (rng.rand(n) * h).astype(np.float32)
    k["octave"] = rng.randint(0, 8, n); k["angle"] = rng.rand(n) * 360; k["class_id"] = -1
    return k


def test_undistort_matches_cv2(oracle):
    k = random_kps(oracle, 20000, 640, 480, 1)
    u = oracle.undistort_keypoints(k, TUM1)
    K = np.array([[TUM1[0], 0, TUM1[2]], [0, TUM1[1], TUM1[3]], [0, 0, 1]], np.float32)
    ref = cv2.undistortPoints(np.stack([k["x"], k["y"]], 1).reshape(-1, 1, 2), K, TUM1[4:], None, K).reshape(-1, 2)
    assert np.array_equal(u["x"].view(np.uint32), ref[:, 0].view(np.uint32))
    assert np.array_equal(u["y"].view(np.uint32), ref[:, 1].view(np.uint32))
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(u[f], k[f])
    assert oracle.undistort_keypoints(k, NODIST).tobytes() == k.tobytes()   # k1 == 0: mvKeysUn = mvKeys


def test_image_bounds(oracle):
    K = np.array([[TUM1[0], 0, TUM1[2]], [0, TUM1[1], TUM1[3]], [0, 0, 1]], np.float32)
    corners = np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32).reshape(-1, 1, 2)
    m = cv2.undistortPoints(corners, K, TUM1[4:], None, K).reshape(-1, 2)
    ref = np.array([min(m[0, 0], m[2, 0]), max(m[1, 0], m[3, 0]), min(m[0, 1], m[1, 1]), max(m[2, 1], m[3, 1])], np.float32)
    assert np.array_equal(oracle.image_bounds(TUM1, 640, 480), ref)
    assert np.array_equal(oracle.image_bounds(NODIST, 1241, 376), np.array([0, 1241, 0, 376], np.float32))


def c_round(v):
    return int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


def py_grid(k, b):
    f32 = np.float32
    invw, invh = f32(64) / (b[1] - b[0]), f32(48) / (b[3] - b[2])
    grid = [[[] for _ in range(48)] for _ in range(64)]
    for i in range(len(k)):
        gx = c_round(float((f32(k["x"][i]) - b[0]) * invw)); gy = c_round(float((f32(k["y"][i]) - b[2]) * invh))
        if 0 <= gx < 64 and 0 <= gy < 48:
            grid[gx][gy].append(i)
    return grid, invw, invh


def py_area(k, b, grid, invw, invh, x, y, r, lo, hi):
    f32 = np.float32
    x, y, r = f32(x), f32(y), f32(r)
    out = []
    cx0 = max(0, int(math.floor(float((x - b[0] - r) * invw))))
    if cx0 >= 64: return out
    cx1 = min(63, int(math.ceil(float((x - b[0] + r) * invw))))
    if cx1 < 0: return out
    cy0 = max(0, int(math.floor(float((y - b[2] - r) * invh))))
    if cy0 >= 48: return out
    cy1 = min(47, int(math.ceil(float((y - b[2] + r) * invh))))
    if cy1 < 0: return out
    check = lo > 0 or hi >= 0
    for ix in range(cx0, cx1 + 1):
        for iy in range(cy0, cy1 + 1):
            for i in grid[ix][iy]:
                if check:
                    if k["octave"][i] < lo: continue
                    if hi >= 0 and k["octave"][i] > hi: continue
                dx = f32(k["x"][i]) - x; dy = f32(k["y"][i]) - y
                if f32(f32(dx * dx) + f32(dy * dy)) < f32(r * r):
                    out.append(i)
    return out


def test_grid_and_area_queries(oracle):
    k = oracle.undistort_keypoints(random_kps(oracle, 3000, 640, 480, 2), TUM1)
    b = oracle.image_bounds(TUM1, 640, 480)
    start, items = oracle.assign_grid(k, b)
    grid, invw, invh = py_grid(k, b)
    flat = [i for ix in range(64) for iy in range(48) for i in grid[ix][iy]]
    assert items.tolist() == flat
    assert [start[ix * 48 + iy + 1] - start[ix * 48 + iy] for ix in range(64) for iy in range(48)] == \
        [len(grid[ix][iy]) for ix in range(64) for iy in range(48)]
    rng = np.random.RandomState(3)
    for _ in range(200):
        x, y, r = rng.rand() * 700 - 30, rng.rand() * 540 - 30, rng.rand() * 120 + 1
        lo, hi = [(-1, -1), (0, 0), (2, 5), (3, -1), (0, 7)][rng.randint(5)]
        assert oracle.features_in_area(k, b, x, y, r, lo, hi).tolist() == py_area(k, b, grid, invw, invh, x, y, r, lo, hi)
