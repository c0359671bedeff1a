"""ctypes binding of oracle/_ref/liborbref.so. TEST INFRASTRUCTURE ONLY.

liborbref.so is the REFERENCE's own ORB_SLAM2::ORBextractor (src/ORBextractor.cc compiled unmodified
from /root/reference by `make -C oracle ref`) behind the C wrapper oracle/ref_wrap.cpp, with the OpenCV
stand-in of oracle/cvshim/ underneath (primitive arithmetic = the oracle's cv2-pinned restatement).
It pins the oracle's CONTROL FLOW: cell loop, iniTh/minTh fallback, DistributeOctTree (with the real
heap-address tie-breaking), IC_Angle, computeOrbDescriptor, level scaling.

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
