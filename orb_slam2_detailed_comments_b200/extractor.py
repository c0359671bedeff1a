"""Host-side mirror of the reference's ORBextractor (include/ORBextractor.h:93-162) over the C ABI.

Names and argument meaning follow the reference: ORBextractor(nfeatures, scaleFactor, nlevels,
iniThFAST, minThFAST); calling the object is operator()(image, mask, keypoints, descriptors);
GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
GetInverseScaleSigmaSquares; mvImagePyramid after a call. Batch entry points are additions
for throughput runs (frames resident in HBM).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import stream_arg, KP_DTYPE, OrbParams, check, lib, ptr


class ORBextractor:
    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0, max_batch=64):
        self._L = lib()
        self._h = C.c_void_p()
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(scaleFactor), int(nlevels)
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)
        self.device = device
        p = OrbParams(self.nfeatures, self.scaleFactor, self.nlevels, self.iniThFAST, self.minThFAST)
        check(self._L.orb_create(C.byref(p), device, max_batch, C.byref(self._h)))
        n = self.nlevels
        self._scale = np.zeros(n, np.float32); self._inv = np.zeros(n, np.float32)
        self._s2 = np.zeros(n, np.float32); self._is2 = np.zeros(n, np.float32)
        self.mnFeaturesPerLevel = np.zeros(n, np.int32)
        check(self._L.orb_get_scale_tables(self._h, ptr(self._scale), ptr(self._inv), ptr(self._s2), ptr(self._is2),
                                           ptr(self.mnFeaturesPerLevel)))
        self.max_keypoints = int(self._L.orb_max_keypoints(self._h))
        self.mvImagePyramid = []

    def max_keypoints_for(self, width, height):
        """Exact output capacity for this image size (per level max(nfeatures_l + 3, 4 * quadtree roots) + 1)."""
        return int(self._L.orb_max_keypoints_for_size(self._h, int(width), int(height))) or self.max_keypoints

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters of the reference (include/ORBextractor.h:119-159)
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv.copy()
    def GetScaleSigmaSquares(self): return self._s2.copy()
    def GetInverseScaleSigmaSquares(self): return self._is2.copy()

    # ---- operator()(image, mask, keypoints, descriptors), src/ORBextractor.cc:1533
    def __call__(self, image, mask=None, want_pyramid=False):
        """Returns (keypoints[KP_DTYPE], descriptors[N,32] u8). Empty image -> (None, None),
        the reference's "outputs untouched". `mask` is ignored, as in the reference."""
        if image is None or image.size == 0:
            return None, None
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected (ORBextractor.cc:1543)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        key = (w, h)
        if getattr(self, "_call_key", None) != key:   # output staging of the drop-in call, reused while the size stays
            cap = self.max_keypoints_for(w, h)
            self._call_key, self._call_cap = key, cap
            self._call_kps, self._call_desc = np.empty(cap, KP_DTYPE), np.empty((cap, 32), np.uint8)
            self._call_n = C.c_int(0)
            self._call_ptrs = (ptr(self._call_kps), ptr(self._call_desc), C.byref(self._call_n))
        cap, kps, desc, n = self._call_cap, self._call_kps, self._call_desc, self._call_n
        n.value = 0
        views = (_lib.OrbLevelView * self.nlevels)() if want_pyramid else None
        check(self._L.orb_extract(self._h, ptr(image), w, h, image.strides[0], self._call_ptrs[0], cap, self._call_ptrs[2], self._call_ptrs[1],
                                  C.cast(views, C.c_void_p) if want_pyramid else None))
        if want_pyramid:
            self.mvImagePyramid = []
            for v in views:
                # interior view with the 19-px border readable around it, like cv::Mat ROI views
                total = (C.c_uint8 * (v.step * (v.height + 38))).from_address(v.data - 19 * v.step - 32)
                full = np.frombuffer(total, np.uint8).reshape(v.height + 38, v.step)
                self.mvImagePyramid.append(full[19:19 + v.height, 32:32 + v.width])
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch_host(self, images):
        """images: (B,H,W) u8 numpy (pinned or pageable). H2D/D2H inside. Returns (kps[B,cap], desc[B,cap,32], counts[B])."""
        assert images.dtype == np.uint8 and images.ndim == 3 and images.flags.c_contiguous
        B, h, w = images.shape
        cap = self.max_keypoints_for(w, h)
        kps = np.zeros((B, cap), KP_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        counts = np.zeros(B, np.int32)
        check(self._L.orb_extract_batch_host(self._h, ptr(images), B, w, h, w, w * h, ptr(kps), cap, ptr(counts), ptr(desc)))
        return kps, desc, counts

    def extract_batch_host_into(self, images, kps, desc, counts, wait=True):
        """Same with caller-provided (e.g. pinned) output arrays: no allocation in the timed path. wait=False returns
        once everything is enqueued (orb_extract_batch_host_async): call synchronize() before reading the outputs or
        reusing the buffers; consecutive calls then form one continuous upload / compute / download pipeline."""
        B, h, w = images.shape
        fn = self._L.orb_extract_batch_host if wait else self._L.orb_extract_batch_host_async
        check(fn(self._h, ptr(images), B, w, h, w, w * h, ptr(kps), kps.shape[1], ptr(counts), ptr(desc)))

    def extract_batch_device(self, d_images, d_kps, d_desc, d_counts, stream=None):
        """Device-resident batch: torch uint8 tensors. d_images (B,H,W); d_kps (B,cap,28) u8;
        d_desc (B,cap,32) u8; d_counts (B) int32. Asynchronous on `stream` (a raw cudaStream_t int)."""
        B, h, w = d_images.shape
        cap = d_kps.shape[1]
        check(self._L.orb_extract_batch_device(self._h, ptr(d_images), B, w, h, d_images.stride(1), d_images.stride(0),
                                               ptr(d_kps), cap, ptr(d_counts), ptr(d_desc), self._stream(stream, d_images)))

    # ---- stereo: both eyes + Frame::ComputeStereoMatches (src/Frame.cc:121-158, 831-1082)
    def extract_stereo(self, left, right, mbf, mb):
        """Returns (kpsL, descL, kpsR, descR, mvuRight, mvDepth) for one rectified pair."""
        assert left.shape == right.shape and left.dtype == np.uint8 and right.dtype == np.uint8
        left = np.ascontiguousarray(left); right = np.ascontiguousarray(right)
        h, w = left.shape
        cap = self.max_keypoints_for(w, h)
        kl = np.zeros(cap, KP_DTYPE); kr = np.zeros(cap, KP_DTYPE)
        dl = np.zeros((cap, 32), np.uint8); dr = np.zeros((cap, 32), np.uint8)
        ur = np.zeros(cap, np.float32); dp = np.zeros(cap, np.float32)
        nl = C.c_int(0); nr = C.c_int(0)
        check(self._L.orb_extract_stereo(self._h, ptr(left), ptr(right), w, h, w, C.c_float(mbf), C.c_float(mb), ptr(kl), cap,
                                         C.byref(nl), ptr(dl), ptr(kr), C.byref(nr), ptr(dr), ptr(ur), ptr(dp)))
        a, b = nl.value, nr.value
        return kl[:a].copy(), dl[:a].copy(), kr[:b].copy(), dr[:b].copy(), ur[:a].copy(), dp[:a].copy()

    def extract_stereo_batch_device(self, d_images, d_kps, d_desc, d_counts, d_uright, d_depth, mbf, mb, stream=None):
        """d_images (2P,H,W) u8 interleaved L0,R0,...; d_uright / d_depth (P,cap) f32."""
        B, h, w = d_images.shape
        cap = d_kps.shape[1]
        check(self._L.orb_extract_stereo_batch_device(self._h, ptr(d_images), B // 2, w, h, d_images.stride(1), d_images.stride(0),
                                                      ptr(d_kps), cap, ptr(d_counts), ptr(d_desc), C.c_float(mbf), C.c_float(mb),
                                                      ptr(d_uright), ptr(d_depth), self._stream(stream, d_images)))

    def _stream(self, stream, tensor):
        # default = torch's current stream of the operands (see _lib.stream_arg), remembered for synchronize()
        self._last_stream = stream_arg(stream, tensor)
        return self._last_stream

    def synchronize(self, stream=None):
        """Waits for `stream`; without one, for everything the handle has in flight and the stream of the last device
        call. Raises on a capacity overflow (ORB_ERR_CAPACITY)."""
        last = getattr(self, "_last_stream", None)
        if not stream and last is not None and last.value:
            check(self._L.orb_synchronize(self._h, last))
        check(self._L.orb_synchronize(self._h, C.c_void_p(stream or 0)))

    def last_launch_count(self):
        return int(self._L.orb_last_launch_count(self._h))

    def last_call_breakdown(self):
        """Host wall-clock microseconds of the last __call__: staging copy, enqueue, device wait, copy-out."""
        us = np.zeros(4, np.float64)
        check(self._L.orb_last_call_breakdown(self._h, ptr(us)))
        return dict(zip(("stage_copy_us", "enqueue_us", "device_wait_us", "copy_out_us"), (float(v) for v in us)))

    STAGES = ("pyramid", "fast", "quadtree", "blur", "describe")

    def set_lanes(self, lanes):
        check(self._L.orb_set_lanes(self._h, int(lanes)))

    def set_profiling(self, enable):
        check(self._L.orb_set_profiling(self._h, int(bool(enable))))

    def stage_times(self):
        """{stage: (total_ms, launches)} accumulated since the previous call (CUDA events on the stream)."""
        ms = np.zeros(5, np.float64); ln = np.zeros(5, np.int64)
        check(self._L.orb_get_stage_times(self._h, ptr(ms), ptr(ln)))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(self.STAGES)}

    # ---- stage outputs of the last call (parity tests)
    def stage_level(self, frame, level):
        w = C.c_int(); h = C.c_int()
        check(self._L.orb_stage_level_size(self._h, level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value + 38, w.value + 38), np.uint8)
        check(self._L.orb_stage_copy_level(self._h, frame, level, ptr(out)))
        return out

    def stage_blur(self, frame, level):
        w = C.c_int(); h = C.c_int()
        check(self._L.orb_stage_level_size(self._h, level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        check(self._L.orb_stage_copy_blur(self._h, frame, level, ptr(out)))
        return out

    def _stage_list(self, fn, frame, level):
        cap = 1 << 20
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = C.c_int(0)
        check(fn(self._h, frame, level, ptr(xs), ptr(ys), ptr(sc), cap, C.byref(n)))
        m = min(n.value, cap)
        return xs[:m].copy(), ys[:m].copy(), sc[:m].copy()

    def stage_candidates(self, frame, level):
        return self._stage_list(self._L.orb_stage_copy_candidates, frame, level)

    def stage_kept(self, frame, level):
        return self._stage_list(self._L.orb_stage_copy_kept, frame, level)
