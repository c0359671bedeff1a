"""Seeded synthetic 8-bit frames and descriptor sets (SURVEY.md §8d "Synthetic inputs").

numpy only, deterministic for a given (w, h, seed): a low-frequency random field, ~400
alpha-blended random rectangles (corners and edges for FAST), Gaussian noise sigma 3.
Used by tests and bench.py; this is workload generation, not part of the hot path.
"""
import numpy as np


def _upsample_bilinear(low, h, w):
    lh, lw = low.shape
    ys = np.linspace(0, lh - 1, h)
    xs = np.linspace(0, lw - 1, w)
    y0 = np.floor(ys).astype(np.int64); y1 = np.minimum(y0 + 1, lh - 1); fy = (ys - y0)[:, None]
    x0 = np.floor(xs).astype(np.int64); x1 = np.minimum(x0 + 1, lw - 1); fx = (xs - x0)[None, :]
    a = low[y0][:, x0]; b = low[y0][:, x1]; c = low[y1][:, x0]; d = low[y1][:, x1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def synth_frame(w, h, seed, n_rect=400, noise_sigma=3.0):
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    low = rng.rand(h // 16 + 2, w // 16 + 2) * 160.0 + 40.0
    img = _upsample_bilinear(low, h, w)
    for _ in range(n_rect):
        rw = int(rng.randint(6, max(8, w // 6))); rh = int(rng.randint(6, max(8, h // 5)))
        x0 = int(rng.randint(-rw // 2, w)); y0 = int(rng.randint(-rh // 2, h))
        val = rng.rand() * 255.0; alpha = 0.35 + 0.65 * rng.rand()
        ys = slice(max(y0, 0), min(y0 + rh, h)); xs = slice(max(x0, 0), min(x0 + rw, w))
        img[ys, xs] = img[ys, xs] * (1 - alpha) + val * alpha
    if noise_sigma > 0:
        img = img + rng.normal(0.0, noise_sigma, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_batch(w, h, count, seed0=0, unique=None):
    """(count, h, w) u8. `unique` < count repeats a pool of unique frames (documented by callers)."""
    unique = count if unique is None else min(unique, count)
    pool = np.stack([synth_frame(w, h, seed0 + i) for i in range(unique)])
    if unique == count:
        return pool
    reps = (count + unique - 1) // unique
    return np.concatenate([pool] * reps)[:count].copy()


def adversarial_frames(w, h):
    """Edge-case frames of SURVEY.md §8d: constant, low contrast, checkerboard, uniform noise."""
    rng = np.random.RandomState(1234)
    out = {}
    out["constant"] = np.full((h, w), 128, np.uint8)
    base = synth_frame(w, h, 99, noise_sigma=0.0).astype(np.float32)
    out["low_contrast"] = np.clip(np.rint(128 + (base - 128) * 0.12), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    out["checkerboard"] = (((yy // 8 + xx // 8) % 2) * 200 + 20).astype(np.uint8)
    out["uniform_noise"] = rng.randint(0, 256, (h, w)).astype(np.uint8)
    return out


def random_descriptors(n, seed):
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    return rng.randint(0, 256, (n, 32)).astype(np.uint8)


def correlated_descriptor_pair(n, seed, flip_prob=0.06, outlier_frac=0.3):
    """Two descriptor sets with realistic match structure: B is a shuffled copy of A with
    per-bit noise, a fraction replaced by unrelated descriptors. Returns (A, B, angA, angB)."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    A = rng.randint(0, 256, (n, 32)).astype(np.uint8)
    bits = np.unpackbits(A, axis=1)
    flips = (rng.rand(*bits.shape) < flip_prob).astype(np.uint8)
    B = np.packbits(bits ^ flips, axis=1)
    perm = rng.permutation(n)
    B = B[perm]
    out = rng.rand(n) < outlier_frac
    B[out] = rng.randint(0, 256, (int(out.sum()), 32)).astype(np.uint8)
    angA = (rng.rand(n) * 360.0).astype(np.float32)
    inv = np.empty(n, np.int64); inv[np.arange(n)] = perm
    angB = ((angA[perm] - 12.0 + rng.normal(0, 3.0, n)) % 360.0).astype(np.float32)
    return A, B, angA, angB


def tracking_scene(n_cur, n_last, seed, w=1241, h=376, fx=718.856, cx=607.1928, cy=185.2157, bf=386.1448,
                   noise_px=3.0, flip_bits=24, frac_mapped=0.8, frac_unobserved=0.05, frac_occupied=0.05, nlevels=8,
                   distinct=0.0):
    """Synthetic state for the projection / BoW matchers: a current frame of n_cur keypoints and a last
    frame of n_last keypoints whose map points project close to current keypoints (position noise
    `noise_px`, descriptor = current descriptor with up to `flip_bits` flipped bits), so that windows
    hold several candidates (`distinct` = fraction of map points that follow a keypoint of their own), some map points compete for one keypoint and ties occur. numpy only."""
    from ._lib import KP_DTYPE
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    f32 = np.float32
    cur = np.zeros(n_cur, KP_DTYPE)
    cur["x"] = (16 + rng.rand(n_cur) * (w - 32)).astype(f32); cur["y"] = (16 + rng.rand(n_cur) * (h - 32)).astype(f32)
    cur["octave"] = np.minimum(rng.geometric(0.35, n_cur) - 1, nlevels - 1); cur["angle"] = (rng.rand(n_cur) * 360).astype(f32)
    cur["size"] = 31; cur["class_id"] = -1
    cur_desc = rng.randint(0, 256, (n_cur, 32)).astype(np.uint8)
    depth_cur = (4 + rng.rand(n_cur) * 40).astype(f32)
    uright = np.where(rng.rand(n_cur) < 0.7, cur["x"] - f32(bf) / depth_cur, -1).astype(f32)
    occupied0 = (rng.rand(n_cur) < frac_occupied).astype(np.uint8)
    # pose: small rotation about y, translation mostly along z
    a = 0.02 * rng.randn()
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([0.05 * rng.randn(), 0.02 * rng.randn(), -0.6 + 0.1 * rng.randn()])
    Tcw = np.eye(4, dtype=f32); Tcw[:3, :3] = R.astype(f32); Tcw[:3, 3] = t.astype(f32)
    # last frame: every keypoint follows one current keypoint (several may follow the same one)
    src = rng.randint(0, n_cur, n_last)
    if distinct > 0:   # tracking-like: most map points follow their own keypoint, a few compete
        own = np.resize(rng.permutation(n_cur), n_last)
        src = np.where(rng.rand(n_last) < distinct, own, src)
    last = np.zeros(n_last, KP_DTYPE)
    last["octave"] = np.clip(cur["octave"][src] + rng.randint(-1, 2, n_last), 0, nlevels - 1)
    rot = 10.0 * rng.randn() + rng.randn(n_last) * 3.0
    wild = rng.rand(n_last) < 0.1
    last["angle"] = np.mod(cur["angle"][src] + rot + wild * rng.rand(n_last) * 360, 360).astype(f32)
    last["size"] = 31; last["class_id"] = -1
    u = cur["x"][src] + rng.randn(n_last) * noise_px; v = cur["y"][src] + rng.randn(n_last) * noise_px
    z = depth_cur[src].astype(np.float64) * (1 + 0.01 * rng.randn(n_last))
    behind = rng.rand(n_last) < 0.02
    z = np.where(behind, -z, z)
    Xc = np.stack([(u - cx) * z / fx, (v - cy) * z / fx, z], 1)
    Xw = ((Xc - t[None, :]) @ R).astype(f32)           # R^T (Xc - t)
    last["x"] = u.astype(f32); last["y"] = v.astype(f32)
    mp_desc = cur_desc[src].copy()
    nflip = rng.randint(0, flip_bits + 1, n_last)
    for i in range(n_last):
        bits = rng.randint(0, 256, nflip[i])
        np.bitwise_xor.at(mp_desc[i], bits >> 3, (1 << (bits & 7)).astype(np.uint8))
    dup = rng.rand(n_last) < 0.05                      # exact duplicates: distance ties between candidates
    mp_desc[dup] = cur_desc[src[dup]]
    flags = (rng.rand(n_last) < frac_mapped).astype(np.uint8)
    flags |= ((rng.rand(n_last) >= frac_unobserved).astype(np.uint8) << 1)
    return {"cur": cur, "cur_desc": cur_desc, "uright": uright, "occupied0": occupied0, "Tcw": Tcw, "last": last, "Xw": Xw,
            "mp_desc": mp_desc, "mp_flags": flags, "cam4": np.array([fx, fx, cx, cy], f32),
            "bounds": np.array([0, w, 0, h], f32), "mbf": f32(bf), "mb": f32(bf / fx), "src": src}


def triangulation_pair(sc, seed):
    """A keyframe pair for SearchForTriangulation from a tracking_scene: keyframe 1 = the scene's last frame (at the
    identity pose), keyframe 2 = its current frame (pose Tcw). Returns the fundamental matrix F12 (x1^T F12 x2 = 0),
    the epipole of camera 1 in image 2 and random map-point / stereo flags."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    f32 = np.float32
    fx, fy, cx, cy = [float(v) for v in sc["cam4"]]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    R = sc["Tcw"][:3, :3].astype(np.float64); t = sc["Tcw"][:3, 3].astype(np.float64)
    # X2 = R X1 + t  ->  E21 = [t]x R with x2^T E21 x1 = 0;  F12 = (K^-T E21 K^-1)^T
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Kinv = np.linalg.inv(K)
    F21 = Kinv.T @ (tx @ R) @ Kinv
    F12 = (F21.T / np.abs(F21).max()).astype(f32)
    e = K @ t                                        # camera centre 1 (origin) seen from camera 2
    n1, n2 = len(sc["last"]), len(sc["cur"])
    kps1 = sc["last"].copy()                         # keyframe 1 sees the world points from the identity pose
    X = sc["Xw"].astype(np.float64)
    z = np.where(np.abs(X[:, 2]) > 1e-3, X[:, 2], 1.0)
    kps1["x"] = (fx * X[:, 0] / z + cx).astype(f32); kps1["y"] = (fy * X[:, 1] / z + cy).astype(f32)
    return {"kps1": kps1, "F12": F12.reshape(9), "ex": f32(e[0] / e[2]), "ey": f32(e[1] / e[2]),
            "has_mp1": (rng.rand(n1) < 0.3).astype(np.uint8), "has_mp2": (rng.rand(n2) < 0.3).astype(np.uint8),
            "ur1": np.where(rng.rand(n1) < 0.5, sc["last"]["x"] - 5, -1).astype(f32)}
