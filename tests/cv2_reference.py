"""cv2-driven restatement of the reference's ORBextractor control flow. TEST INFRASTRUCTURE.

The closest thing to the real reference that runs in this image: the reference's control
flow (ORBextractor.cc:1037-1184, 1533-1724) re-expressed in Python, delegating every pixel
primitive to the real OpenCV (python cv2 4.13.0): cv2.resize, cv2.copyMakeBorder,
cv2.FastFeatureDetector per cell, cv2.GaussianBlur, cv2.fastAtan2. The quadtree is a
Python restatement with the canonical tie rule (later-created node first among equal
sizes). Used to pin the C++ oracle end-to-end and to mint tests/golden/*.npz.
"""
import math

import cv2
import numpy as np

EDGE = 19
HALF_PATCH = 15


def load_pattern():
    """The 256x4 int8 sampling pattern from include/orb_pattern_data.h (a DATA table)."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "orb_pattern_data.h")
    txt = open(path).read()
    body = txt[txt.index("{") + 1:txt.index("};")]
    nums = [int(v) for v in re.findall(r"-?\d+", body)]
    assert len(nums) == 1024
    return np.asarray(nums, np.int32)


def c_round_half_even(v):
    return int(np.rint(np.float32(v)))


class Cv2Reference:
    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th, pattern):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = float(np.float32(scale_factor))  # double member initialised from a float argument
        self.scale = [np.float32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(np.float32(float(self.scale[-1]) * sf))
        self.inv_scale = [np.float32(1.0) / s for s in self.scale]
        factor = np.float32(1.0 / sf)
        want = np.float32(nfeatures) * (np.float32(1) - factor) / (np.float32(1) - np.float32(math.pow(float(factor), float(nlevels))))
        self.per_level = []
        for _ in range(nlevels - 1):
            self.per_level.append(c_round_half_even(want))
            want = np.float32(want * factor)
        self.per_level.append(max(nfeatures - sum(self.per_level), 0))
        umax = [0] * (HALF_PATCH + 1)
        vmax = int(math.floor(np.float32(HALF_PATCH) * np.sqrt(np.float32(2.0)) / 2 + 1))
        vmin = int(math.ceil(np.float32(HALF_PATCH) * np.sqrt(np.float32(2.0)) / 2))
        for v in range(vmax + 1):
            umax[v] = int(np.rint(math.sqrt(HALF_PATCH * HALF_PATCH - v * v)))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        self.pattern = np.asarray(pattern, np.int32).reshape(512, 2)
        self.fast_ini = cv2.FastFeatureDetector_create(ini_th, True)
        self.fast_min = cv2.FastFeatureDetector_create(min_th, True)

    def pyramid(self, image):
        levels = []
        for l in range(self.nlevels):
            s = self.inv_scale[l]
            w = c_round_half_even(np.float32(image.shape[1]) * s)
            h = c_round_half_even(np.float32(image.shape[0]) * s)
            if l == 0:
                inner = image
            else:
                prev = levels[-1][EDGE:-EDGE, EDGE:-EDGE]
                inner = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
            levels.append(cv2.copyMakeBorder(inner, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))
        return levels

    def detect_level(self, bordered, N):
        """Returns list of (x_rel, y_rel, response) in candidate order + kept indices."""
        img = bordered[EDGE:-EDGE, EDGE:-EDGE]
        H, W = img.shape
        # coordinates below are in level-image coords; mvImagePyramid[level] is the interior view,
        # whose rows/cols extend (negative / beyond) into the border memory.
        minB = EDGE - 3
        maxBX, maxBY = W - EDGE + 3, H - EDGE + 3
        width, height = float(maxBX - minB), float(maxBY - minB)
        ncols, nrows = int(width / 30), int(height / 30)
        wcell, hcell = int(math.ceil(width / ncols)), int(math.ceil(height / nrows))
        cand = []
        stats = dict(cells=0, fallback=0)
        for i in range(nrows):
            iniY = minB + i * hcell
            maxY = iniY + hcell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(ncols):
                iniX = minB + j * wcell
                maxX = iniX + wcell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                sub = bordered[EDGE + iniY:EDGE + maxY, EDGE + iniX:EDGE + maxX]
                stats["cells"] += 1
                kps = self.fast_ini.detect(sub)
                if len(kps) == 0:
                    kps = self.fast_min.detect(sub)
                    if len(kps):
                        stats["fallback"] += 1
                for kp in kps:
                    cand.append((kp.pt[0] + j * wcell, kp.pt[1] + i * hcell, kp.response))
        kept = distribute_quadtree(cand, minB, maxBX, minB, maxBY, N) if cand else []
        return cand, kept, stats

    def ic_angle(self, bordered, x, y):
        cx, cy = int(np.rint(x)) + EDGE, int(np.rint(y)) + EDGE
        m01 = m10 = 0
        for v in range(-HALF_PATCH, HALF_PATCH + 1):
            d = self.umax[abs(v)]
            row = bordered[cy + v, cx - d:cx + d + 1].astype(np.int64)
            us = np.arange(-d, d + 1)
            m10 += int((us * row).sum())
            m01 += v * int(row.sum())
        return float(cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10))))

    def descriptor(self, blurred, x, y, angle):
        ang = np.float32(angle) * np.float32(math.pi / 180.0)  # factorPI = (float)(CV_PI/180.f)
        a = np.float32(math.cos(float(ang))); b = np.float32(math.sin(float(ang)))
        px = self.pattern[:, 0].astype(np.float32); py = self.pattern[:, 1].astype(np.float32)
        ry = np.rint(px * b + py * a).astype(np.int64)  # float32 arithmetic, round half even
        rx = np.rint(px * a - py * b).astype(np.int64)
        vals = blurred[int(np.rint(y)) + ry, int(np.rint(x)) + rx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)

    def __call__(self, image):
        levels = self.pyramid(image)
        out_kps, out_desc, dbg = [], [], []
        per_level = []
        for l, bordered in enumerate(levels):
            cand, kept, stats = self.detect_level(bordered, self.per_level[l])
            dbg.append(dict(cand=cand, kept=kept, **stats))
            patch = int(np.float32(31) * self.scale[l])
            kps = []
            for k in kept:
                x = cand[k][0] + (EDGE - 3); y = cand[k][1] + (EDGE - 3)
                kps.append([x, y, float(patch), self.ic_angle(bordered, x, y), cand[k][2], l])
            per_level.append(kps)
        for l, kps in enumerate(per_level):
            if not kps:
                continue
            inner = levels[l][EDGE:-EDGE, EDGE:-EDGE].copy()
            blurred = cv2.GaussianBlur(inner, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            dbg[l]["blur"] = blurred
            for kp in kps:
                out_desc.append(self.descriptor(blurred, kp[0], kp[1], kp[3]))
                if l != 0:
                    kp[0] = float(np.float32(kp[0]) * self.scale[l]); kp[1] = float(np.float32(kp[1]) * self.scale[l])
                out_kps.append(kp)
        return out_kps, (np.stack(out_desc) if out_desc else np.zeros((0, 32), np.uint8)), levels, dbg


class _Node:
    __slots__ = ("x0", "x1", "y0", "y1", "keys", "frozen", "seq")


def _divide(n, cand):
    hx = int(math.ceil(np.float32(n.x1 - n.x0) / 2)); hy = int(math.ceil(np.float32(n.y1 - n.y0) / 2))
    mx, my = n.x0 + hx, n.y0 + hy
    ch = []
    for (x0, x1, y0, y1) in ((n.x0, mx, n.y0, my), (mx, n.x1, n.y0, my), (n.x0, mx, my, n.y1), (mx, n.x1, my, n.y1)):
        c = _Node(); c.x0, c.x1, c.y0, c.y1 = x0, x1, y0, y1; c.keys = []; c.frozen = False; c.seq = 0
        ch.append(c)
    for k in n.keys:
        x, y = cand[k][0], cand[k][1]
        if x < mx:
            ch[0 if y < my else 2].keys.append(k)
        else:
            ch[1 if y < my else 3].keys.append(k)
    for c in ch:
        c.frozen = len(c.keys) == 1
    return ch


def distribute_quadtree(cand, minX, maxX, minY, maxY, N):
    """Python restatement of DistributeOctTree (ORBextractor.cc:688-1033), canonical tie rule.
    `nodes` is a Python list used as the std::list (index 0 = front)."""
    nIni = int(round(float(np.float32(maxX - minX) / np.float32(maxY - minY))))
    hX = np.float32(maxX - minX) / np.float32(nIni)
    nodes = []
    for i in range(nIni):
        r = _Node(); r.x0 = int(hX * np.float32(i)); r.x1 = int(hX * np.float32(i + 1)); r.y0 = 0; r.y1 = maxY - minY
        r.keys = []; r.frozen = False; r.seq = -i - 1
        nodes.append(r)
    roots = list(nodes)
    for k, c in enumerate(cand):
        roots[int(np.float32(c[0]) / hX)].keys.append(k)
    nodes = [n for n in nodes if n.keys]
    for n in nodes:
        n.frozen = len(n.keys) == 1
    counter = [0]

    def emit(children, expandable):
        cnt = 0
        for c in children:
            if not c.keys:
                continue
            c.seq = counter[0]; counter[0] += 1
            nodes.insert(0, c)
            if len(c.keys) > 1:
                cnt += 1
                expandable.append(c)
        return cnt

    done = False
    while not done:
        prev = len(nodes)
        expandable = []
        n_to_expand = 0
        for n in list(nodes):  # snapshot = front-to-back walk that never revisits pushed-front children
            if n.frozen:
                continue
            n_to_expand += emit(_divide(n, cand), expandable)
            nodes.remove(n)
        if len(nodes) >= N or len(nodes) == prev:
            done = True
        elif len(nodes) + 3 * n_to_expand > N:
            while not done:
                prev = len(nodes)
                todo = sorted(expandable, key=lambda n: (len(n.keys), n.seq))
                expandable = []
                for n in reversed(todo):
                    emit(_divide(n, cand), expandable)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    done = True
    kept = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if cand[k][2] > cand[best][2]:
                best = k
        kept.append(best)
    return kept
