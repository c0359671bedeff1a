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
