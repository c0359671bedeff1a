"""Host-side mirror of the reference's ORBmatcher hot path (include/ORBmatcher.h:57-221) over the C ABI.

ORBmatcher(nnratio=0.6, checkOri=True); DescriptorDistance(a, b) (static in the reference);
SearchForInitialization(F1, F2, vbPrevMatched, windowSize=10) -> (nmatches, vnMatches12), with
F1/F2 anything exposing mvKeysUn-like arrays (see FrameView). TH_LOW / TH_HIGH / HISTO_LENGTH
as in src/ORBmatcher.cc:49-51.
"""
import ctypes as C

import numpy as np

from ._lib import stream_arg, OrbFrameView, OrbMatchParams, check, lib, ptr


class FrameView:
    """The slice of Frame that SearchForInitialization reads: undistorted keypoint positions,
    octaves, angles, descriptors, and the image bounds mnMinX..mnMaxY used by the 64x48 grid."""

    def __init__(self, xy, octave, angle, descriptors, bounds):
        self.xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        self.octave = np.ascontiguousarray(octave, np.int32)
        self.angle = np.ascontiguousarray(angle, np.float32)
        self.descriptors = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        self.bounds = tuple(float(b) for b in bounds)  # (mnMinX, mnMaxX, mnMinY, mnMaxY)
        self.N = len(self.angle)

    @staticmethod
    def from_keypoints(kps, descriptors, width, height):
        return FrameView(np.stack([kps["x"], kps["y"]], 1), kps["octave"], kps["angle"], descriptors,
                         (0.0, float(width), 0.0, float(height)))

    def c_view(self):
        return OrbFrameView(self.N, ptr(self.xy), ptr(self.octave), ptr(self.angle), ptr(self.descriptors))


class ORBmatcher:
    TH_LOW = 50
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio=0.6, checkOri=True, device=0, max_keypoints=4096, max_pairs=1):
        self._L = lib()
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self._h = C.c_void_p()
        self.device = device
        check(self._L.orb_matcher_create(device, max_pairs, max_keypoints, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orb_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def DescriptorDistance(a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        assert a.size == 32 and b.size == 32
        return int(lib().orb_descriptor_distance(ptr(a), ptr(b)))

    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=10, mode=0, want_distances=False):
        """mode 0: reference-faithful windowed search; mode 1: brute force over all of F2.
        vbPrevMatched (N1 x 2 float32) is updated in place, as in the reference."""
        prev = np.ascontiguousarray(vbPrevMatched, np.float32).reshape(-1, 2)
        m12 = np.full(F1.N, -1, np.int32)
        best = np.zeros(F1.N, np.int32); second = np.zeros(F1.N, np.int32)
        nm = C.c_int(0)
        mp = OrbMatchParams(self.mfNNratio, int(self.mbCheckOrientation), int(windowSize), int(mode),
                            F2.bounds[0], F2.bounds[1], F2.bounds[2], F2.bounds[3])
        v1, v2 = F1.c_view(), F2.c_view()
        check(self._L.orb_search_for_initialization(self._h, C.byref(v1), C.byref(v2), C.byref(mp), ptr(prev), ptr(m12),
                                                    C.byref(nm), ptr(best), ptr(second)))
        if isinstance(vbPrevMatched, np.ndarray) and vbPrevMatched.dtype == np.float32:
            vbPrevMatched.reshape(-1, 2)[:] = prev
        if want_distances:
            return nm.value, m12, best, second
        return nm.value, m12

    # ---- throughput entry points on device-resident tensors (torch)
    def match_pairs_device(self, d_desc, d_angle, d_matches12, d_nmatches, stream=None):
        """d_desc (2P,n,32) u8, d_angle (2P,n) f32, d_matches12 (P,n) i32, d_nmatches (P) i32."""
        P = d_desc.shape[0] // 2
        n = d_desc.shape[1]
        check(self._L.orb_match_pairs_device(self._h, ptr(d_desc), ptr(d_angle), P, n, self.mfNNratio,
                                             int(self.mbCheckOrientation), ptr(d_matches12), ptr(d_nmatches),
                                             self._stream(stream, d_desc)))

    def match_allpairs_device(self, d_all, row_begin, row_end, d_counts, stream=None, col_begin=0, col_end=None):
        """d_all (nKF,nDesc,32) u8 holding ALL keyframes; d_counts (row_end-row_begin, nKF) i32.
        Only columns [col_begin, col_end) are computed by this call."""
        nkf, nd = d_all.shape[0], d_all.shape[1]
        col_end = nkf if col_end is None else col_end
        check(self._L.orb_match_allpairs_device(self._h, ptr(d_all), nkf, nd, row_begin, row_end, col_begin, col_end,
                                                self.mfNNratio, ptr(d_counts), self._stream(stream, d_all)))

    def hamming_matrix_device(self, d_a, d_b, d_out, stream=None):
        check(self._L.orb_hamming_matrix_device(self._h, ptr(d_a), d_a.shape[0], ptr(d_b), d_b.shape[0], ptr(d_out),
                                                self._stream(stream, d_a)))

    def _stream(self, stream, tensor):
        # default = torch's current stream of the operands (see _lib.stream_arg), remembered for synchronize()
        self._last_stream = stream_arg(stream, tensor)
        return self._last_stream

    def synchronize(self, stream=None):
        """Waits for `stream`; without one, for the handle's own stream and the stream of the last device call."""
        check(self._L.orb_matcher_synchronize(self._h, C.c_void_p(stream or 0)))
        last = getattr(self, "_last_stream", None)
        if not stream and last is not None and last.value:
            check(self._L.orb_matcher_synchronize(self._h, last))


def int_pipe_peak(device=0):
    """Measured POPC and LOP3 issue rates (ops/s) of the device: matching roofline denominators."""
    out = {}
    for name, what in (("popc", 0), ("lop3", 1), ("mix_popc_4lop3", 2)):
        v = C.c_double(0)
        check(lib().orb_int_pipe_peak(device, what, C.byref(v)))
        out[name] = v.value
    return out
