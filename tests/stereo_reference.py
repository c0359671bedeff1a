"""Python/cv2 restatement of Frame::ComputeStereoMatches (src/Frame.cc:831-1082). TEST INFRASTRUCTURE.
Independent of the C++ oracle: patches via numpy slices, L1 norm via cv2.norm as in the reference."""
import math

import cv2
import numpy as np

f32 = np.float32


def c_round(v):
    return f32(math.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))


def compute_stereo_matches(levels_l, levels_r, scale, inv_scale, kps_l, desc_l, kps_r, desc_r, mbf, mb):
    """levels_*: list of bordered pyramid levels ((h+38) x (w+38) arrays)."""
    N, Nr = len(kps_l), len(kps_r)
    u_right = np.full(N, -1.0, np.float32); depth = np.full(N, -1.0, np.float32)
    th_orb = (100 + 50) // 2
    n_rows = levels_l[0].shape[0] - 38
    rows = [[] for _ in range(n_rows)]
    for i in range(Nr):
        y = f32(kps_r["y"][i]); r = f32(2.0) * f32(scale[kps_r["octave"][i]])
        for yi in range(int(math.floor(float(y - r))), int(math.ceil(float(y + r))) + 1):
            rows[yi].append(i)
    mbf, mb = f32(mbf), f32(mb)
    max_d = mbf / mb
    D = np.unpackbits(desc_l[:, None, :] ^ desc_r[None, :, :], axis=2).sum(2) if N and Nr else np.zeros((N, Nr), int)
    dist_idx = []
    for il in range(N):
        lvl = int(kps_l["octave"][il]); vl = f32(kps_l["y"][il]); ul = f32(kps_l["x"][il])
        cands = rows[int(vl)]
        if not cands:
            continue
        min_u, max_u = ul - max_d, ul
        if max_u < 0:
            continue
        best, best_r = 100, 0
        for ir in cands:
            o = int(kps_r["octave"][ir])
            if o < lvl - 1 or o > lvl + 1:
                continue
            ur = f32(kps_r["x"][ir])
            if min_u <= ur <= max_u:
                d = int(D[il, ir])
                if d < best:
                    best, best_r = d, ir
        if best >= th_orb:
            continue
        sf = f32(inv_scale[lvl])
        su = c_round(ul * sf); sv = c_round(vl * sf); sr = c_round(f32(kps_r["x"][best_r]) * sf)
        w = L = 5
        PL = levels_l[lvl]; PR = levels_r[lvl]
        cu, cv_, cr = int(su) + 19, int(sv) + 19, int(sr) + 19
        IL = PL[cv_ - w:cv_ + w + 1, cu - w:cu + w + 1].astype(np.float32)
        IL = IL - IL[w, w]
        if sr - L - w < 0 or sr + L + w + 1 >= PR.shape[1] - 38:
            continue
        best_d, best_inc = 2 ** 31 - 1, 0
        vd = [f32(0)] * (2 * L + 1)
        for inc in range(-L, L + 1):
            IR = PR[cv_ - w:cv_ + w + 1, cr + inc - w:cr + inc + w + 1].astype(np.float32)
            IR = IR - IR[w, w]
            dist = f32(cv2.norm(IL, IR, cv2.NORM_L1))
            if dist < best_d:
                best_d, best_inc = int(dist), inc
            vd[L + inc] = dist
        if best_inc in (-L, L):
            continue
        d1, d2, d3 = vd[L + best_inc - 1], vd[L + best_inc], vd[L + best_inc + 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = (d1 - d3) / (f32(2.0) * (d1 + d3 - f32(2.0) * d2))
        if delta < -1 or delta > 1:
            continue
        best_ur = f32(scale[lvl]) * (f32(sr) + f32(best_inc) + delta)
        disp = ul - best_ur
        if disp >= 0 and disp < max_d:
            if disp <= 0:
                disp = f32(0.01); best_ur = f32(float(ul) - 0.01)
            depth[il] = mbf / disp
            u_right[il] = best_ur
            dist_idx.append((best_d, il))
    if dist_idx:
        dist_idx.sort()
        median = f32(dist_idx[len(dist_idx) // 2][0])
        th = f32(1.5) * f32(1.4) * median
        for d, il in reversed(dist_idx):
            if f32(d) < th:
                break
            u_right[il] = -1; depth[il] = -1
    return u_right, depth
