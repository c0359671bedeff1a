"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module. The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborb_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_ic_angle.restype = C.c_float
        L.orc_extract_batch_mt.restype = C.c_long
        L.orc_match_batch_mt.restype = C.c_long
        _LIB = L
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


class OracleExtractor:
    """Mirror of ORBextractor (include/ORBextractor.h:93-162) on the CPU oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.L.orc_create(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th))
        sc = np.zeros(nlevels, np.float32); isc = sc.copy(); s2 = sc.copy(); is2 = sc.copy()
        per = np.zeros(nlevels, np.int32); umax = np.zeros(16, np.int32)
        self.L.orc_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(per), _p(umax))
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = sc, isc, s2, is2
        self.per_level, self.umax = per, umax

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 64 * self.nlevels + 1024
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_extract(self.h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap)
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        """Bordered pyramid level l as (h+38, w+38) array."""
        w = C.c_int(); h = C.c_int(); s = C.c_int()
        self.L.orc_level_info(self.h, l, C.byref(w), C.byref(h), C.byref(s))
        out = np.zeros((h.value + 38, w.value + 38), np.uint8)
        self.L.orc_level_copy(self.h, l, _p(out))
        return out

    def blurred(self, l):
        w = C.c_int(); h = C.c_int(); s = C.c_int()
        self.L.orc_level_info(self.h, l, C.byref(w), C.byref(h), C.byref(s))
        out = np.zeros((h.value, w.value), np.uint8)
        ok = self.L.orc_blur_copy(self.h, l, _p(out))
        return out if ok else None

    def candidates(self, l):
        cap = 1 << 20
        xs = np.zeros(cap, np.int32); ys = xs.copy(); sc = xs.copy()
        n = self.L.orc_level_candidates(self.h, l, _p(xs), _p(ys), _p(sc), cap)
        assert n <= cap
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()

    def kept(self, l):
        cap = 1 << 16
        idx = np.zeros(cap, np.int32)
        n = self.L.orc_level_kept(self.h, l, _p(idx), cap)
        return idx[:n].copy()

    def stats(self, l):
        a = C.c_int(); b = C.c_int(); c = C.c_int()
        self.L.orc_level_stats(self.h, l, C.byref(a), C.byref(b), C.byref(c))
        return dict(cells=a.value, fallback=b.value, tie_sensitive=c.value)


def stereo_matches(ext_left, ext_right, kps_l, desc_l, kps_r, desc_r, mbf, mb):
    """Mirror of Frame::ComputeStereoMatches (Frame.cc:831) on the pyramids the two oracle extractors
    hold from their last call. Returns (mvuRight, mvDepth, n_kept)."""
    kps_l = np.ascontiguousarray(kps_l); kps_r = np.ascontiguousarray(kps_r)
    desc_l = np.ascontiguousarray(desc_l, np.uint8); desc_r = np.ascontiguousarray(desc_r, np.uint8)
    ur = np.zeros(len(kps_l), np.float32); dp = np.zeros(len(kps_l), np.float32)
    n = lib().orc_stereo_matches(ext_left.h, ext_right.h, _p(kps_l), len(kps_l), _p(desc_l), _p(kps_r), len(kps_r),
                                 _p(desc_r), C.c_float(mbf), C.c_float(mb), _p(ur), _p(dp))
    return ur, dp, n


def undistort_keypoints(kps, cam9):
    """Frame::UndistortKeyPoints (Frame.cc:724): cam9 = (fx, fy, cx, cy, k1, k2, p1, p2, k3)."""
    kps = np.ascontiguousarray(kps); cam = np.ascontiguousarray(cam9, np.float32)
    out = np.zeros(len(kps), KP_DTYPE)
    lib().orc_undistort_keypoints(_p(kps), len(kps), _p(cam), _p(out))
    return out


def image_bounds(cam9, w, h):
    cam = np.ascontiguousarray(cam9, np.float32); b = np.zeros(4, np.float32)
    lib().orc_image_bounds(_p(cam), w, h, _p(b))
    return b


def assign_grid(kps, bounds):
    kps = np.ascontiguousarray(kps); b = np.ascontiguousarray(bounds, np.float32)
    start = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(max(len(kps), 1), np.int32)
    lib().orc_assign_grid(_p(kps), len(kps), _p(b), _p(start), _p(items))
    return start, items[:start[-1]].copy()


def features_in_area(kps, bounds, x, y, r, min_level=-1, max_level=-1):
    kps = np.ascontiguousarray(kps); b = np.ascontiguousarray(bounds, np.float32)
    out = np.zeros(max(len(kps), 1), np.int32)
    n = lib().orc_features_in_area(_p(kps), len(kps), _p(b), C.c_float(x), C.c_float(y), C.c_float(r), min_level,
                                   max_level, _p(out), len(out))
    return out[:n].copy()


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def border101(src):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((h + 38, w + 38), np.uint8)
    lib().orc_border101(_p(src), w, h, src.strides[0], _p(dst))
    return dst


def fast(img, threshold, nms=True):
    """img may be a non-contiguous 2-D view (row stride honoured)."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    h, w = img.shape
    cap = max(w * h, 1)
    xs = np.zeros(cap, np.int32); ys = xs.copy(); sc = xs.copy()
    n = lib().orc_fast(C.c_void_p(img.ctypes.data), w, h, img.strides[0], threshold, int(nms), _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def fast_score_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h, w), np.int32)
    lib().orc_fast_score_map(_p(img), w, h, img.strides[0], _p(out))
    return out


def gauss7(src):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((h, w), np.uint8)
    lib().orc_gauss7(_p(src), w, h, src.strides[0], _p(dst), w)
    return dst


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(y.shape, np.float32)
    lib().orc_fast_atan2_many(_p(y), _p(x), _p(out), y.size)
    return out


def brief(img, px, py, angle):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orc_brief(C.c_void_p(img.ctypes.data), img.strides[0], C.c_float(px), C.c_float(py), C.c_float(angle), _p(d))
    return d


def quadtree(xs, ys, score, min_x, max_x, min_y, max_y, N):
    xs = np.ascontiguousarray(xs, np.float32); ys = np.ascontiguousarray(ys, np.float32)
    score = np.ascontiguousarray(score, np.int32)
    kept = np.zeros(len(xs) + 8, np.int32)
    tie = C.c_int(0)
    n = lib().orc_quadtree(_p(xs), _p(ys), _p(score), len(xs), min_x, max_x, min_y, max_y, N, _p(kept), len(kept), C.byref(tie))
    return kept[:n].copy(), tie.value


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming(_p(a), _p(b))


def hamming_matrix(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    out = np.zeros((a.shape[0], b.shape[0]), np.int32)
    lib().orc_hamming_matrix(_p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


def three_maxima(histo):
    histo = np.ascontiguousarray(histo, np.int32)
    out = np.zeros(3, np.int32)
    lib().orc_three_maxima(_p(histo), len(histo), _p(out))
    return tuple(int(v) for v in out)


def search_for_initialization(xy1, oct1, ang1, desc1, xy2, oct2, ang2, desc2, bounds, prev_matched,
                              window=100, nnratio=0.9, check_ori=True, mode=0):
    """Mirror of ORBmatcher::SearchForInitialization (ORBmatcher.cc:573).

    mode 0 = reference-faithful windowed search, mode 1 = brute force over all of F2.
    Returns (nmatches, matches12, prev_matched_updated, best, second)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    xy1, xy2, ang1, ang2 = f32(xy1), f32(xy2), f32(ang1), f32(ang2)
    oct1, oct2 = i32(oct1), i32(oct2)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    n1, n2 = len(ang1), len(ang2)
    prev = f32(prev_matched).copy()
    b = f32(bounds)
    m12 = np.full(n1, -1, np.int32)
    best = np.zeros(n1, np.int32); second = np.zeros(n1, np.int32)
    n = lib().orc_search_for_initialization(n1, _p(xy1), _p(oct1), _p(ang1), _p(desc1), n2, _p(xy2), _p(oct2),
                                            _p(ang2), _p(desc2), _p(b), _p(prev), _p(m12), int(window),
                                            C.c_float(nnratio), int(check_ori), int(mode), _p(best), _p(second))
    return n, m12, prev, best, second


PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("angle", "<f4"),
                             ("min_level", "<i4"), ("max_level", "<i4"), ("flags", "<i4")])
SEARCH_BEST, SEARCH_RATIO_LEVEL, SEARCH_RATIO = 0, 1, 2


def search_by_projection(kps_un, desc, uright, bounds, occupied0, queries, qdesc, mode, th, ratio=0.0, check_ori=False):
    """ORBmatcher::SearchByProjection on prepared queries (ORBmatcher.cc:72 local map = SEARCH_RATIO_LEVEL,
    :1710 last frame = SEARCH_BEST). -> (nmatches, match_of_keypoint, match_of_query)."""
    kps_un = np.ascontiguousarray(kps_un); desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE); qdesc = np.ascontiguousarray(qdesc, np.uint8)
    n, nq = len(kps_un), len(queries)
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    occ = np.ascontiguousarray(occupied0, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32)
    mk = np.full(max(n, 1), -1, np.int32); mq = np.full(max(nq, 1), -1, np.int32)
    nm = lib().orc_search_by_projection(_p(kps_un), n, _p(desc), _p(ur) if ur is not None else None, _p(b), _p(occ), _p(queries),
                                        _p(qdesc), nq, int(mode), int(th), C.c_float(ratio), int(check_ori), _p(mk), _p(mq))
    return nm, mk[:n], mq[:nq]


def project_last_frame(Xw, mp_flags, kps1, Tcw, cam4, bounds, mbf, th, scale_factors, direction):
    """Projection part of SearchByProjection(CurrentFrame, LastFrame, ...), ORBmatcher.cc:1734-1775 -> queries."""
    Xw = np.ascontiguousarray(Xw, np.float32); fl = np.ascontiguousarray(mp_flags, np.uint8); kps1 = np.ascontiguousarray(kps1)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16); cam = np.ascontiguousarray(cam4, np.float32)
    b = np.ascontiguousarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    out = np.zeros(len(kps1), PROJ_QUERY_DTYPE)
    lib().orc_project_last_frame(len(kps1), _p(Xw), _p(fl), _p(kps1), _p(T), _p(cam), _p(b), C.c_float(mbf), C.c_float(th), _p(sf),
                                 int(direction), _p(out))
    return out


def search_by_bow(kps1, desc1, node1, usable1, kps2, desc2, node2, th=50, ratio=0.7, check_ori=True, unusable2=None):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...), ORBmatcher.cc:247. -> (nmatches, match_of_keypoint (over frame 2,
    values = keyframe feature indices), match_of_query (over keyframe features))."""
    kps1 = np.ascontiguousarray(kps1); kps2 = np.ascontiguousarray(kps2)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    node1 = np.ascontiguousarray(node1, np.int32); node2 = np.ascontiguousarray(node2, np.int32)
    us = np.ascontiguousarray(usable1, np.uint8)
    n1, n2 = len(kps1), len(kps2)
    mk = np.full(max(n2, 1), -1, np.int32); mq = np.full(max(n1, 1), -1, np.int32)
    un2 = None if unusable2 is None else np.ascontiguousarray(unusable2, np.uint8)
    nm = lib().orc_search_by_bow(_p(kps1), n1, _p(desc1), _p(node1), _p(us), _p(kps2), n2, _p(desc2), _p(node2),
                                 _p(un2) if un2 is not None else None, int(th),
                                 C.c_float(ratio), int(check_ori), _p(mk), _p(mq))
    return nm, mk[:n2], mq[:n1]


TRI_PAIR_DTYPE = np.dtype([("F12", "<f4", (9,)), ("ex", "<f4"), ("ey", "<f4"), ("only_stereo", "<i4")])


def search_for_triangulation(kps1, desc1, node1, has_mp1, ur1, kps2, desc2, node2, has_mp2, ur2, pair, scale_factors, level_sigma2,
                             check_ori=True):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:884). -> (nmatches, matches12)."""
    c = lambda a, t: np.ascontiguousarray(a, t)
    kps1, kps2 = np.ascontiguousarray(kps1), np.ascontiguousarray(kps2)
    desc1, desc2 = c(desc1, np.uint8), c(desc2, np.uint8)
    node1, node2, has_mp1, has_mp2 = c(node1, np.int32), c(node2, np.int32), c(has_mp1, np.uint8), c(has_mp2, np.uint8)
    u1 = None if ur1 is None else c(ur1, np.float32); u2 = None if ur2 is None else c(ur2, np.float32)
    pair = np.ascontiguousarray(pair, TRI_PAIR_DTYPE).reshape(1)
    sf, s2 = c(scale_factors, np.float32), c(level_sigma2, np.float32)
    m12 = np.full(max(len(kps1), 1), -1, np.int32)
    nm = lib().orc_search_for_triangulation(_p(kps1), len(kps1), _p(desc1), _p(node1), _p(has_mp1), _p(u1) if u1 is not None else None,
                                            _p(kps2), len(kps2), _p(desc2), _p(node2), _p(has_mp2), _p(u2) if u2 is not None else None,
                                            _p(pair), _p(sf), _p(s2), int(check_ori), _p(m12))
    return nm, m12[:len(kps1)]


def cvt_gray(img, blue_first=False):
    """cv::cvtColor(img, CV_RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY) (Tracking.cc:250-276)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, c = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().orc_cvt_gray(_p(img), w, h, C.c_size_t(img.strides[0]), c, int(blue_first), _p(out))
    return out


def remap_linear(img, mapx, mapy):
    """cv::remap(img, ., mapx, mapy, INTER_LINEAR), float maps, BORDER_CONSTANT 0 (stereo_euroc.cc:181-188)."""
    img = np.ascontiguousarray(img, np.uint8); mapx = np.ascontiguousarray(mapx, np.float32); mapy = np.ascontiguousarray(mapy, np.float32)
    dh, dw = mapx.shape
    out = np.zeros((dh, dw), np.uint8)
    lib().orc_remap_linear(_p(img), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]), _p(mapx), _p(mapy), dw, dh, _p(out))
    return out


def distinctive_descriptors(desc, offsets):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:365) for map points with observations desc[offsets[p]:offsets[p+1]]."""
    desc = np.ascontiguousarray(desc, np.uint8); offsets = np.ascontiguousarray(offsets, np.int32)
    out = np.zeros(len(offsets) - 1, np.int32)
    lib().orc_distinctive_descriptors(_p(desc), _p(offsets), len(out), _p(out))
    return out


def allpairs_counts(desc, nnratio=0.9, row_begin=0, row_end=None):
    desc = np.ascontiguousarray(desc, np.uint8)
    nkf, nd, _ = desc.shape
    row_end = nkf if row_end is None else row_end
    out = np.zeros((row_end - row_begin, nkf), np.int32)
    lib().orc_allpairs_counts(_p(desc), nkf, nd, C.c_float(nnratio), row_begin, row_end, _p(out))
    return out


def extract_batch_mt(imgs, nfeatures, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, nthreads=None):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    B, h, w = imgs.shape
    nthreads = nthreads or hardware_threads()
    counts = np.zeros(B, np.int32)
    total = lib().orc_extract_batch_mt(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th, _p(imgs), B, w, h,
                                       nthreads, _p(counts))
    return total, counts


def match_batch_mt(desc, angle, nnratio=0.9, nthreads=None):
    """desc: (2P, n, 32) u8, angle: (2P, n) f32 -> brute-force SearchForInitialization per pair."""
    desc = np.ascontiguousarray(desc, np.uint8); angle = np.ascontiguousarray(angle, np.float32)
    P = desc.shape[0] // 2; n = desc.shape[1]
    nthreads = nthreads or hardware_threads()
    nm = np.zeros(P, np.int32)
    total = lib().orc_match_batch_mt(_p(desc), _p(angle), P, n, C.c_float(nnratio), nthreads, _p(nm))
    return total, nm


def hardware_threads():
    return int(lib().orc_hardware_threads())
