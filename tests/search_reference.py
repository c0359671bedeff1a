"""Second, independent restatement of the ordered matchers in plain Python (lists, dicts, numpy float32
scalars), written from ORBmatcher.cc:72-169 (local map), :1710-1860 (last frame), :247-420 (BoW) and
Frame.cc:590-670. The C++ reference needs OpenCV/DBoW2 and cannot be built here, so the oracle's
matchers are pinned against this restatement, cv2.gemm (projection arithmetic) and cv2.norm (Hamming)."""
import math

import numpy as np

f32 = np.float32
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming(a, b):
    return int(POP[np.bitwise_xor(a, b)].sum())


def c_round(v):
    return int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


class PyFrame:
    def __init__(self, kps, desc, bounds, uright=None):
        self.k, self.desc, self.b, self.uright = kps, desc, [f32(x) for x in bounds], uright
        self.invw = f32(64) / (self.b[1] - self.b[0]); self.invh = f32(48) / (self.b[3] - self.b[2])
        self.grid = {}
        for i in range(len(kps)):
            gx = c_round(f32(f32(kps["x"][i] - self.b[0]) * self.invw)); gy = c_round(f32(f32(kps["y"][i] - self.b[2]) * self.invh))
            if 0 <= gx < 64 and 0 <= gy < 48:
                self.grid.setdefault((gx, gy), []).append(i)

    def area(self, x, y, r, min_level=-1, max_level=-1):
        x, y, r = f32(x), f32(y), f32(r)
        out = []
        cx0 = max(0, int(math.floor(f32(f32(f32(x - self.b[0]) - r) * self.invw))))
        if cx0 >= 64: return out
        cx1 = min(63, int(math.ceil(f32(f32(f32(x - self.b[0]) + r) * self.invw))))
        if cx1 < 0: return out
        cy0 = max(0, int(math.floor(f32(f32(f32(y - self.b[2]) - r) * self.invh))))
        if cy0 >= 48: return out
        cy1 = min(47, int(math.ceil(f32(f32(f32(y - self.b[2]) + r) * self.invh))))
        if cy1 < 0: return out
        check = min_level > 0 or max_level >= 0
        for ix in range(cx0, cx1 + 1):
            for iy in range(cy0, cy1 + 1):
                for i in self.grid.get((ix, iy), ()):
                    o = int(self.k["octave"][i])
                    if check:
                        if o < min_level: continue
                        if max_level >= 0 and o > max_level: continue
                    dx = f32(self.k["x"][i] - x); dy = f32(self.k["y"][i] - y)
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(r * r):
                        out.append(i)
        return out


def three_maxima(cnt):
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(cnt):
        if s > m1: m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2: m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3: m3, i3 = s, i
    if m2 < f32(0.1) * f32(m1): i2 = i3 = -1
    elif m3 < f32(0.1) * f32(m1): i3 = -1
    return i1, i2, i3


def rot_bin(a1, a2):
    rot = f32(a1) - f32(a2)
    if rot < 0: rot = f32(rot + f32(360))
    b = c_round(f32(rot * f32(30 / f32(360.0))))
    return 0 if b == 30 else b


def finish(hist, match_kp, nm):
    a, b, c = three_maxima([len(h) for h in hist])
    for i in range(30):
        if i in (a, b, c): continue
        for idx in hist[i]:
            match_kp[idx] = -1; nm -= 1
    return nm


def search_local_map(F, occupied0, mps, th_factor, scale_factors, nnratio, TH_HIGH=100):
    """mps: list of dicts(track_in_view, bad, level, view_cos, x, y, xr, desc, nobs). ORBmatcher.cc:72-169."""
    occ = [bool(o) for o in occupied0]
    match_kp = [-1] * len(F.k); nm = 0
    for q, mp in enumerate(mps):
        if not mp["track_in_view"] or mp["bad"]: continue
        r = f32(2.5) if f32(mp["view_cos"]) > f32(0.998) else f32(4.0)
        if th_factor != 1.0: r = f32(r * f32(th_factor))
        rad = f32(r * f32(scale_factors[mp["level"]]))
        cands = F.area(mp["x"], mp["y"], rad, mp["level"] - 1, mp["level"])
        if not cands: continue
        bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
        for i in cands:
            if occ[i]: continue
            if F.uright[i] > 0:
                if abs(f32(f32(mp["xr"]) - F.uright[i])) > rad: continue
            d = hamming(mp["desc"], F.desc[i])
            if d < bd: bd2, bd, bl2, bl, bi = bd, d, bl, int(F.k["octave"][i]), i
            elif d < bd2: bl2, bd2 = int(F.k["octave"][i]), d
        if bd <= TH_HIGH:
            if bl == bl2 and f32(bd) > f32(f32(nnratio) * f32(bd2)): continue
            match_kp[bi] = q; nm += 1
            if mp["nobs"] > 0: occ[bi] = True
    return nm, match_kp


def search_last_frame(F, occupied0, last_kps, Xw, mp_flags, mp_desc, Tcw, cam4, mbf, th, scale_factors, direction,
                      check_ori=True, TH_HIGH=100):
    """ORBmatcher.cc:1710-1860; direction 0 / 1 (bForward) / 2 (bBackward) decided by the caller."""
    import cv2
    occ = [bool(o) for o in occupied0]
    match_kp = [-1] * len(F.k); nm = 0
    hist = [[] for _ in range(30)]
    R = np.ascontiguousarray(Tcw[:3, :3], f32); t = np.ascontiguousarray(Tcw[:3, 3:4], f32)
    fx, fy, cx, cy = [f32(c) for c in cam4]
    for i in range(len(last_kps)):
        if not (mp_flags[i] & 1): continue
        xc3 = cv2.gemm(R, np.ascontiguousarray(Xw[i].reshape(3, 1), f32), 1.0, t, 1.0)
        xc, yc = f32(xc3[0, 0]), f32(xc3[1, 0])
        with np.errstate(divide="ignore"):
            invz = f32(np.float64(1.0) / np.float64(xc3[2, 0]))
        if invz < 0: continue
        u = f32(f32(f32(fx * xc) * invz) + cx); v = f32(f32(f32(fy * yc) * invz) + cy)
        if u < F.b[0] or u > F.b[1] or v < F.b[2] or v > F.b[3]: continue
        o = int(last_kps["octave"][i])
        rad = f32(f32(th) * f32(scale_factors[o]))
        if direction == 1: cands = F.area(u, v, rad, o)
        elif direction == 2: cands = F.area(u, v, rad, 0, o)
        else: cands = F.area(u, v, rad, o - 1, o + 1)
        if not cands: continue
        bd, bi = 256, -1
        for j in cands:
            if occ[j]: continue
            if F.uright[j] > 0:
                ur = f32(u - f32(f32(mbf) * invz))
                if abs(f32(ur - F.uright[j])) > rad: continue
            d = hamming(mp_desc[i], F.desc[j])
            if d < bd: bd, bi = d, j
        if bd <= TH_HIGH:
            match_kp[bi] = i; nm += 1
            if mp_flags[i] & 2: occ[bi] = True
            if check_ori: hist[rot_bin(last_kps["angle"][i], F.k["angle"][bi])].append(bi)
    if check_ori: nm = finish(hist, match_kp, nm)
    return nm, match_kp


def search_by_bow(k1, d1, node1, usable1, k2, d2, node2, nnratio=0.7, check_ori=True, TH_LOW=50, usable2=None, strict=False):
    """ORBmatcher.cc:247-420 with FeatureVectors rebuilt from per-feature node ids; usable2 + strict = the
    keyframe-keyframe variant :729-880 (candidates need a good map point, bestDist1 < TH_LOW)."""
    fv1, fv2 = {}, {}
    for i, n in enumerate(node1):
        if n >= 0: fv1.setdefault(int(n), []).append(i)
    for i, n in enumerate(node2):
        if n >= 0: fv2.setdefault(int(n), []).append(i)
    match = [-1] * len(k2); nm = 0
    hist = [[] for _ in range(30)]
    for node in sorted(set(fv1) & set(fv2)):
        for i1 in fv1[node]:
            if not usable1[i1]: continue
            b1, b2, bi = 256, 256, -1
            for i2 in fv2[node]:
                if match[i2] >= 0: continue
                if usable2 is not None and not usable2[i2]: continue
                d = hamming(d1[i1], d2[i2])
                if d < b1: b2, b1, bi = b1, d, i2
                elif d < b2: b2 = d
            if (b1 < TH_LOW if strict else b1 <= TH_LOW) and f32(b1) < f32(f32(nnratio) * f32(b2)):
                match[bi] = i1; nm += 1
                if check_ori: hist[rot_bin(k1["angle"][i1], k2["angle"][bi])].append(bi)
    if check_ori: nm = finish(hist, match, nm)
    return nm, match


def search_for_triangulation(k1, d1, node1, has_mp1, ur1, k2, d2, node2, has_mp2, ur2, F12, ex, ey, only_stereo, sf, sigma2,
                             check_ori=True, TH_LOW=50):
    """ORBmatcher.cc:884-1100 + CheckDistEpipolarLine :205-227."""
    fv1, fv2 = {}, {}
    for i, n in enumerate(node1):
        if n >= 0: fv1.setdefault(int(n), []).append(i)
    for i, n in enumerate(node2):
        if n >= 0: fv2.setdefault(int(n), []).append(i)
    F = np.asarray(F12, f32).reshape(3, 3)
    matched2 = [False] * len(k2); m12 = [-1] * len(k1); nm = 0
    hist = [[] for _ in range(30)]
    for node in sorted(set(fv1) & set(fv2)):
        for i1 in fv1[node]:
            if has_mp1[i1]: continue
            st1 = ur1 is not None and ur1[i1] >= 0
            if only_stereo and not st1: continue
            x1, y1 = f32(k1["x"][i1]), f32(k1["y"][i1])
            best, bi = TH_LOW, -1
            for i2 in fv2[node]:
                if matched2[i2] or has_mp2[i2]: continue
                st2 = ur2 is not None and ur2[i2] >= 0
                if only_stereo and not st2: continue
                d = hamming(d1[i1], d2[i2])
                if d > TH_LOW or d > best: continue
                x2, y2, o2 = f32(k2["x"][i2]), f32(k2["y"][i2]), int(k2["octave"][i2])
                if not st1 and not st2:
                    dx, dy = f32(f32(ex) - x2), f32(f32(ey) - y2)
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(f32(100) * f32(sf[o2])): continue
                a = f32(f32(f32(x1 * F[0, 0]) + f32(y1 * F[1, 0])) + F[2, 0])
                b = f32(f32(f32(x1 * F[0, 1]) + f32(y1 * F[1, 1])) + F[2, 1])
                c = f32(f32(f32(x1 * F[0, 2]) + f32(y1 * F[1, 2])) + F[2, 2])
                num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
                den = f32(f32(a * a) + f32(b * b))
                if den == 0: continue
                dsqr = f32(f32(num * num) / den)
                if np.float64(dsqr) < np.float64(3.84) * np.float64(f32(sigma2[o2])):
                    bi, best = i2, d
            if bi >= 0:
                m12[i1] = bi; matched2[bi] = True; nm += 1
                if check_ori: hist[rot_bin(k1["angle"][i1], k2["angle"][bi])].append(i1)
    if check_ori:
        a, b, c = three_maxima([len(h) for h in hist])
        for i in range(30):
            if i in (a, b, c): continue
            for i1 in hist[i]:
                m12[i1] = -1; nm -= 1
    return nm, m12
