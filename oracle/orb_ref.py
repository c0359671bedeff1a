"""ctypes binding of oracle/_ref/liborbref.so. TEST INFRASTRUCTURE ONLY.

liborbref.so is the REFERENCE's own code - src/ORBextractor.cc, ORBmatcher.cc, Frame.cc, MapPoint.cc, KeyFrame.cc,
Map.cc compiled unmodified from /root/reference by `make -C oracle ref` - behind the C wrapper oracle/ref_wrap.cpp,
with the OpenCV stand-in of oracle/cvshim/ underneath (primitive arithmetic = the oracle's cv2-pinned restatement).
It pins the oracle's CONTROL FLOW AND ARITHMETIC: cell loop, iniTh/minTh fallback, DistributeOctTree (with the real
heap-address tie-breaking), IC_Angle, computeOrbDescriptor, level scaling; DescriptorDistance, SearchForInitialization,
the Frame grid / undistortion / area queries, ComputeStereoMatches; the projection / BoW / triangulation searches on
live MapPoint / KeyFrame objects; ComputeDistinctiveDescriptors.

The library is built in the development container (where /root/reference exists) and travels to the
GPU box as a prebuilt, git-ignored binary. `available()` says whether it is there; nothing here ever
reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "liborbref.so")
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def build(reference="/root/reference"):
    """Compile the reference extractor if its sources are present (development container only)."""
    if os.path.exists(os.path.join(reference, "src", "ORBextractor.cc")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REFERENCE=" + reference])
    return available()


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise FileNotFoundError(_PATH + " (build it with `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(_PATH)
        L.orbref_create.restype = C.c_void_p
        L.orbref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbref_destroy.argtypes = [C.c_void_p]
        L.orbref_extract_batch_mt.restype = C.c_long
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def last_call_us():
    """Wall time (us) of the last call into the reference's class made by this module (the matcher / stereo member only,
    measured inside ref_wrap.cpp around the member call)."""
    f = lib().orbref_last_call_us
    f.restype = C.c_double
    return float(f())


class ReferenceExtractor:
    """ORB_SLAM2::ORBextractor of the reference itself (include/ORBextractor.h:93)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.L.orbref_create(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th))
        t = [np.zeros(nlevels, np.float32) for _ in range(4)]
        self.L.orbref_tables(self.h, *[_p(a) for a in t])
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = t

    def __del__(self):
        try:
            self.L.orbref_destroy(self.h)
        except Exception:
            pass

    def __call__(self, img, canonical=True):
        """canonical=True: heap addresses grow with creation order (the tie rule of the oracle and the CUDA
        path, see oracle/ref_wrap.cpp); False: plain malloc, as in the reference's own build."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 64 * self.nlevels + 1024
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orbref_extract(self.h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap, 1 if canonical else 0)
        assert n <= cap
        self.last_n = n
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        """mvImagePyramid[l] with its 19-px border: (h+38, w+38)."""
        w = C.c_int(); h = C.c_int()
        self.L.orbref_level_info(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value + 38, w.value + 38), np.uint8)
        self.L.orbref_level_copy(self.h, l, _p(out))
        return out


def extract_batch_mt(params, imgs, nthreads):
    """One reference extractor per thread over a (B, h, w) stack; returns the total keypoint count."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    B, h, w = imgs.shape
    nf, sf, nl, ini, mn = params
    return lib().orbref_extract_batch_mt(nf, C.c_float(sf), nl, ini, mn, _p(imgs), B, w, h, nthreads)


# ---- ORBmatcher.cc / Frame.cc of the reference (compiled unmodified into the same library) ------------------
def descriptor_distance(a, b):
    """ORBmatcher::DescriptorDistance (ORBmatcher.cc:2083)."""
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orbref_descriptor_distance(_p(a), _p(b))


def matcher_constants():
    a = C.c_int(); b = C.c_int(); c = C.c_int()
    lib().orbref_matcher_constants(C.byref(a), C.byref(b), C.byref(c))
    return dict(TH_LOW=a.value, TH_HIGH=b.value, HISTO_LENGTH=c.value)


def three_maxima(counts):
    """ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2035) on histogram bin sizes."""
    counts = np.ascontiguousarray(counts, np.int32)
    out = np.zeros(3, np.int32)
    lib().orbref_three_maxima(_p(counts), len(counts), _p(out))
    return tuple(int(v) for v in out)


class ReferenceFrame:
    """An ORB_SLAM2::Frame built from given keypoints and descriptors: the reference's own UndistortKeyPoints,
    ComputeImageBounds, AssignFeaturesToGrid and GetFeaturesInArea (Frame.cc:724, :779, :399, :590).
    cam9 = fx fy cx cy k1 k2 p1 p2 k3."""

    def __init__(self, kps, desc, cam9, w, h):
        self.L = lib()
        self.L.orbref_frame_create.restype = C.c_void_p
        self.L.orbref_frame_destroy.argtypes = [C.c_void_p]
        kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        cam9 = np.ascontiguousarray(cam9, np.float32)
        assert len(cam9) == 9 and desc.shape == (len(kps), 32)
        self.n = len(kps)
        self.h = C.c_void_p(self.L.orbref_frame_create(_p(kps), self.n, _p(desc), _p(cam9), int(w), int(h)))

    def __del__(self):
        try:
            self.L.orbref_frame_destroy(self.h)
        except Exception:
            pass

    def keys_un(self):
        out = np.zeros(self.n, KP_DTYPE)
        self.L.orbref_frame_keys_un(self.h, _p(out))
        return out

    def bounds(self):
        b = np.zeros(4, np.float32)
        self.L.orbref_frame_bounds(self.h, _p(b))
        return b

    def grid(self):
        start = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(max(self.n, 1), np.int32)
        self.L.orbref_frame_grid(self.h, _p(start), _p(items))
        return start, items[:start[-1]]

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(self.n, 1), np.int32)
        n = self.L.orbref_frame_features_in_area(self.h, C.c_float(x), C.c_float(y), C.c_float(r), int(min_level), int(max_level),
                                                 _p(out), len(out))
        return out[:n].copy()


def search_for_initialization(F1, F2, prev_matched, window=100, nnratio=0.9, check_ori=True):
    """ORBmatcher(nnratio, check_ori).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, window)
    (ORBmatcher.cc:573). Returns (nmatches, vnMatches12, vbPrevMatched after the call)."""
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.full(F1.n, -1, np.int32)
    n = lib().orbref_search_for_initialization(F1.h, F2.h, _p(prev), _p(m12), int(window), C.c_float(nnratio), int(check_ori))
    return n, m12, prev


def stereo_matches(ref_left, ref_right, mbf, mb):
    """Frame::ComputeStereoMatches (Frame.cc:831) on what the two ReferenceExtractors hold from their last call."""
    n = max(1, ref_left.last_n)
    ur = np.zeros(n, np.float32); dp = np.zeros(n, np.float32)
    kept = lib().orbref_stereo_matches(ref_left.h, ref_right.h, C.c_float(mbf), C.c_float(mb), _p(ur), _p(dp))
    return ur[:ref_left.last_n], dp[:ref_left.last_n], kept


def search_last_frame(cur_frame, uright, occupied0, last_kps, Xw, mp_flags, mp_desc, Tcw, cam4, mbf, mb, th, direction, scale_factors):
    """ORBmatcher(0.9, true).SearchByProjection(CurrentFrame, LastFrame, th, bMono=false) (ORBmatcher.cc:1710) on live
    MapPoint objects. Returns (nmatches, CurrentFrame.mvpMapPoints as indices into the last frame, -1 = none)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    last_kps = np.ascontiguousarray(last_kps, KP_DTYPE)
    uright, Xw, Tcw, cam4, sf = f32(uright), f32(Xw), f32(Tcw), f32(cam4), f32(scale_factors)
    occupied0, mp_flags, mp_desc = u8(occupied0), u8(mp_flags), u8(mp_desc)
    out = np.full(cur_frame.n, -1, np.int32)
    n = lib().orbref_search_last_frame(cur_frame.h, _p(uright), _p(occupied0), _p(last_kps), len(last_kps), _p(Xw), _p(mp_flags),
                                       _p(mp_desc), _p(Tcw), _p(cam4), C.c_float(mbf), C.c_float(mb), C.c_float(th), int(direction),
                                       _p(sf), len(sf), _p(out))
    return n, out


def search_local_map(cur_frame, uright, occupied0, mps, th, nnratio, cam4, mbf, mb, scale_factors):
    """ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:72). mps: list of dicts with the
    MapPoint tracking fields (track_in_view, bad, level, view_cos, x, y, xr, desc, nobs)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    n = len(mps)
    tiv = u8([m["track_in_view"] for m in mps]); bad = u8([m["bad"] for m in mps])
    lvl = np.ascontiguousarray([m["level"] for m in mps], np.int32); nobs = np.ascontiguousarray([m["nobs"] for m in mps], np.int32)
    vc = f32([m["view_cos"] for m in mps]); x = f32([m["x"] for m in mps]); y = f32([m["y"] for m in mps]); xr = f32([m["xr"] for m in mps])
    desc = u8(np.stack([m["desc"] for m in mps]))
    uright, cam4, sf, occupied0 = f32(uright), f32(cam4), f32(scale_factors), u8(occupied0)
    out = np.full(cur_frame.n, -1, np.int32)
    k = lib().orbref_search_local_map(cur_frame.h, _p(uright), _p(occupied0), n, _p(tiv), _p(bad), _p(lvl), _p(vc), _p(x), _p(y), _p(xr),
                                      _p(desc), _p(nobs), C.c_float(th), C.c_float(nnratio), _p(cam4), C.c_float(mbf), C.c_float(mb),
                                      _p(sf), len(sf), _p(out))
    return k, out


def _bow_args(kps, desc, node, usable):
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    node = np.ascontiguousarray(node, np.int32)
    usable = None if usable is None else np.ascontiguousarray(usable, np.uint8)
    return kps, desc, node, usable


def search_by_bow_frame(kps1, desc1, node1, usable1, kps2, desc2, node2, nnratio, check_ori, scale_factors):
    """ORBmatcher(nnratio, check_ori).SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (ORBmatcher.cc:247) on a live KeyFrame.
    Returns (nmatches, per frame feature: the keyframe feature it was matched to or -1)."""
    k1, d1, n1, u1 = _bow_args(kps1, desc1, node1, usable1); k2, d2, n2, _ = _bow_args(kps2, desc2, node2, None)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    out = np.full(len(k2), -1, np.int32)
    n = lib().orbref_search_by_bow_frame(_p(k1), len(k1), _p(d1), _p(n1), _p(u1), _p(k2), len(k2), _p(d2), _p(n2), C.c_float(nnratio),
                                         int(check_ori), _p(sf), len(sf), _p(out))
    return n, out


def search_by_bow_keyframes(kps1, desc1, node1, usable1, kps2, desc2, node2, usable2, nnratio, check_ori, scale_factors):
    """ORBmatcher(nnratio, check_ori).SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (ORBmatcher.cc:729).
    Returns (nmatches, per keyframe-1 feature: the keyframe-2 feature or -1)."""
    k1, d1, n1, u1 = _bow_args(kps1, desc1, node1, usable1); k2, d2, n2, u2 = _bow_args(kps2, desc2, node2, usable2)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    out = np.full(len(k1), -1, np.int32)
    n = lib().orbref_search_by_bow_keyframes(_p(k1), len(k1), _p(d1), _p(n1), _p(u1), _p(k2), len(k2), _p(d2), _p(n2), _p(u2),
                                             C.c_float(nnratio), int(check_ori), _p(sf), len(sf), _p(out))
    return n, out


def search_for_triangulation(kps1, desc1, node1, has_mp1, ur1, kps2, desc2, node2, has_mp2, ur2, Tcw2, cam4, F12, only_stereo,
                             check_ori, scale_factors, level_sigma2):
    """ORBmatcher(0.6, check_ori).SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (ORBmatcher.cc:884),
    keyframe 1 at the identity pose, keyframe 2 at Tcw2. Returns (nmatches, matches12)."""
    k1, d1, n1, h1 = _bow_args(kps1, desc1, node1, has_mp1); k2, d2, n2, h2 = _bow_args(kps2, desc2, node2, has_mp2)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u1 = f32(np.full(len(k1), -1) if ur1 is None else ur1); u2 = f32(np.full(len(k2), -1) if ur2 is None else ur2)
    T, cam4, F12, sf, s2 = f32(Tcw2), f32(cam4), f32(F12), f32(scale_factors), f32(level_sigma2)
    out = np.full(len(k1), -1, np.int32)
    n = lib().orbref_search_for_triangulation(_p(k1), len(k1), _p(d1), _p(n1), _p(h1), _p(u1), _p(k2), len(k2), _p(d2), _p(n2), _p(h2),
                                              _p(u2), _p(T), _p(cam4), _p(F12), int(only_stereo), int(check_ori), _p(sf), _p(s2), len(sf),
                                              _p(out))
    return n, out


def distinctive_descriptors(desc, offsets):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:365) on live MapPoint / KeyFrame objects, for map points with
    observations desc[offsets[p]:offsets[p+1]] (every map point needs >= 1). Returns the chosen descriptor of every point."""
    desc = np.ascontiguousarray(desc, np.uint8); offsets = np.ascontiguousarray(offsets, np.int32)
    out = np.zeros((len(offsets) - 1, 32), np.uint8)
    lib().orbref_distinctive_descriptors(_p(desc), _p(offsets), len(out), _p(out))
    return out
