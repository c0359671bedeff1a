"""ctypes binding of oracle/_ref/liborbref_gpu.so. TEST INFRASTRUCTURE ONLY.

liborbref_gpu.so is the REFERENCE's own src/Frame.cc, ORBmatcher.cc, MapPoint.cc, KeyFrame.cc, Map.cc compiled unmodified
(`make -C oracle refgpu`) and linked against the DROP-IN instead of the reference's src/ORBextractor.cc: the translation units
of orb_slam2_detailed_comments_b200/compat/ provide ORBextractor's constructor and operator(), ORBmatcher::DescriptorDistance /
SearchForInitialization and Frame::ComputeStereoMatches on top of liborb_b200.so (the CUDA kernels). The GPU tests run the
reference's real Frame constructors through it and compare the members they fill with the all-CPU reference (orb_ref.py).
Built in the development container; travels to the GPU box as a prebuilt, git-ignored binary.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "liborbref_gpu.so")
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

EXPORTS = ["orbgpu_last_error", "orbgpu_extractor_create", "orbgpu_extractor_destroy", "orbgpu_extract", "orbgpu_pyramid_level",
           "orbgpu_frame_mono", "orbgpu_frame_stereo", "orbgpu_frame_destroy", "orbgpu_frame_n", "orbgpu_frame_n_right",
           "orbgpu_frame_keys", "orbgpu_frame_descriptors", "orbgpu_frame_stereo_vectors", "orbgpu_frame_bounds", "orbgpu_frame_grid",
           "orbgpu_frame_scale_tables", "orbgpu_search_for_initialization", "orbgpu_descriptor_distance"]


def build(reference="/root/reference"):
    if os.path.exists(os.path.join(reference, "src", "Frame.cc")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "refgpu", "REFERENCE=" + reference])
    return available()


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise FileNotFoundError(_PATH + " (build it with `make -C oracle refgpu` where /root/reference exists)")
        L = C.CDLL(_PATH)
        vp = C.c_void_p
        L.orbgpu_last_error.restype = C.c_char_p
        L.orbgpu_extractor_create.restype = vp
        L.orbgpu_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbgpu_extractor_destroy.argtypes = [vp]
        L.orbgpu_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
        L.orbgpu_pyramid_level.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]
        L.orbgpu_frame_mono.restype = vp
        L.orbgpu_frame_mono.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_float, C.c_float]
        L.orbgpu_frame_stereo.restype = vp
        L.orbgpu_frame_stereo.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_float, C.c_float]
        L.orbgpu_frame_destroy.argtypes = [vp]
        L.orbgpu_frame_n.argtypes = [vp]
        L.orbgpu_frame_n_right.argtypes = [vp]
        L.orbgpu_frame_keys.argtypes = [vp, C.c_int, vp]
        L.orbgpu_frame_descriptors.argtypes = [vp, C.c_int, vp]
        L.orbgpu_frame_stereo_vectors.argtypes = [vp, vp, vp]
        L.orbgpu_frame_bounds.argtypes = [vp]
        L.orbgpu_frame_grid.argtypes = [vp, vp, vp]
        L.orbgpu_frame_scale_tables.argtypes = [vp, C.POINTER(C.c_int), vp, vp, vp, vp]
        L.orbgpu_search_for_initialization.argtypes = [vp, vp, vp, vp, C.c_int, C.c_float, C.c_int]
        L.orbgpu_descriptor_distance.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _err():
    return lib().orbgpu_last_error().decode("utf-8", "replace")


class DropInExtractor:
    """`new ORB_SLAM2::ORBextractor(...)` of the reference's header, body = orb_b200_extractor.cpp (CUDA)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        h = self.L.orbgpu_extractor_create(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th)
        if not h:
            raise RuntimeError(_err())
        self.h = C.c_void_p(h)

    def close(self):
        if self.h:
            self.L.orbgpu_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 64 * self.nlevels + 1024
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orbgpu_extract(self.h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap)
        if n < 0:
            raise RuntimeError(_err())
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        w = C.c_int(); h = C.c_int()
        self.L.orbgpu_pyramid_level(self.h, l, C.byref(w), C.byref(h), None)
        out = np.zeros((h.value + 38, w.value + 38), np.uint8)
        self.L.orbgpu_pyramid_level(self.h, l, C.byref(w), C.byref(h), _p(out))
        return out


class DropInFrame:
    """An ORB_SLAM2::Frame made by the reference's REAL constructors (src/Frame.cc:313 monocular, :121 stereo)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(_err())
        self.L = lib()
        self.h = C.c_void_p(handle)
        self.n = self.L.orbgpu_frame_n(self.h)
        self.n_right = self.L.orbgpu_frame_n_right(self.h)

    @staticmethod
    def mono(extractor, img, cam9, bf=40.0, th_depth=40.0):
        img = np.ascontiguousarray(img, np.uint8); cam9 = np.ascontiguousarray(cam9, np.float32)
        h, w = img.shape
        return DropInFrame(lib().orbgpu_frame_mono(extractor.h, _p(img), w, h, img.strides[0], _p(cam9), C.c_float(bf), C.c_float(th_depth)))

    @staticmethod
    def stereo(ex_left, ex_right, left, right, cam9, bf, th_depth=35.0):
        left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
        cam9 = np.ascontiguousarray(cam9, np.float32)
        h, w = left.shape
        return DropInFrame(lib().orbgpu_frame_stereo(ex_left.h, ex_right.h, _p(left), _p(right), w, h, left.strides[0], _p(cam9),
                                                     C.c_float(bf), C.c_float(th_depth)))

    def close(self):
        if self.h:
            self.L.orbgpu_frame_destroy(self.h)
            self.h = None

    def keys(self, which=0):
        """0 mvKeys, 1 mvKeysUn, 2 mvKeysRight"""
        out = np.zeros(self.n_right if which == 2 else self.n, KP_DTYPE)
        self.L.orbgpu_frame_keys(self.h, which, _p(out))
        return out

    def descriptors(self, right=False):
        out = np.zeros((self.n_right if right else self.n, 32), np.uint8)
        self.L.orbgpu_frame_descriptors(self.h, int(right), _p(out))
        return out

    def stereo_vectors(self):
        ur = np.zeros(max(self.n, 1), np.float32); dp = np.zeros(max(self.n, 1), np.float32)
        self.L.orbgpu_frame_stereo_vectors(self.h, _p(ur), _p(dp))
        return ur[:self.n], dp[:self.n]

    def bounds(self):
        b = np.zeros(4, np.float32)
        self.L.orbgpu_frame_bounds(_p(b))
        return b

    def grid(self):
        start = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(max(self.n, 1), np.int32)
        self.L.orbgpu_frame_grid(self.h, _p(start), _p(items))
        return start, items[:start[-1]]

    def scale_tables(self):
        nl = C.c_int(); t = [np.zeros(16, np.float32) for _ in range(4)]
        self.L.orbgpu_frame_scale_tables(self.h, C.byref(nl), *[_p(a) for a in t])
        return [a[:nl.value] for a in t]


def search_for_initialization(F1, F2, prev_matched, window=100, nnratio=0.9, check_ori=True):
    """ORBmatcher(nnratio, check_ori).SearchForInitialization(F1, F2, ...) through the reference's class, body = CUDA."""
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.full(F1.n, -1, np.int32)
    n = lib().orbgpu_search_for_initialization(F1.h, F2.h, _p(prev), _p(m12), int(window), C.c_float(nnratio), int(check_ori))
    if n < 0:
        raise RuntimeError(_err())
    return n, m12, prev


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orbgpu_descriptor_distance(_p(a), _p(b))


_API = None


def reference_api():
    """The whole oracle/orb_ref.py binding (frames from given keypoints, the searches on live MapPoint / KeyFrame objects ...)
    bound to liborbref_gpu.so instead of liborbref.so: oracle/ref_wrap.cpp is compiled into both libraries, so every orbref_*
    entry point exists in both with the same signature - here the reference's code runs with the drop-in bodies
    (extractor, DescriptorDistance, SearchForInitialization, SearchByProjection(CurrentFrame, LastFrame), ComputeStereoMatches)
    linked in, and its own CPU bodies for everything else."""
    global _API
    if _API is None:
        import importlib.util
        from . import orb_ref
        spec = importlib.util.spec_from_file_location("oracle._orb_ref_on_dropin", orb_ref.__file__)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod._PATH = _PATH
        mod._LIB = None
        _API = mod
    return _API
