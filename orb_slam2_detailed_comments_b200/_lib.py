"""ctypes loader for liborb_b200.so (the C ABI of include/orb_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and
every compute call fails with the CUDA error when no B200 is visible.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liborb_b200.so")

ORB_OK, ORB_ERR_INVALID, ORB_ERR_CUDA, ORB_ERR_CAPACITY, ORB_ERR_UNSUPPORTED = range(5)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class OrbLevelView(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("step", C.c_int64)]


class OrbCamera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("k1", C.c_float),
                ("k2", C.c_float), ("p1", C.c_float), ("p2", C.c_float), ("k3", C.c_float)]


AREA_QUERY_DTYPE = np.dtype([("frame", "<i4"), ("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"),
                             ("max_level", "<i4")])


PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("angle", "<f4"),
                             ("min_level", "<i4"), ("max_level", "<i4"), ("flags", "<i4")])
assert PROJ_QUERY_DTYPE.itemsize == 32
ORB_SEARCH_BEST, ORB_SEARCH_RATIO_LEVEL, ORB_SEARCH_RATIO = 0, 1, 2


class OrbDeviceFrames(C.Structure):
    _fields_ = [("keypoints_un", C.c_void_p), ("descriptors", C.c_void_p), ("uright", C.c_void_p), ("occupied", C.c_void_p),
                ("counts", C.c_void_p), ("cell_start", C.c_void_p), ("cell_items", C.c_void_p), ("bounds", C.c_float * 4),
                ("batch", C.c_int32), ("capacity", C.c_int32)]


TRI_PAIR_DTYPE = np.dtype([("F12", "<f4", (9,)), ("ex", "<f4"), ("ey", "<f4"), ("only_stereo", "<i4")])
assert TRI_PAIR_DTYPE.itemsize == 48


class OrbSearchParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("th", C.c_int32), ("nn_ratio", C.c_float), ("check_orientation", C.c_int32)]


class OrbFrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("xy", C.c_void_p), ("octave", C.c_void_p), ("angle", C.c_void_p),
                ("descriptors", C.c_void_p)]


class OrbMatchParams(C.Structure):
    _fields_ = [("nnratio", C.c_float), ("check_orientation", C.c_int32), ("window", C.c_int32),
                ("mode", C.c_int32), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float),
                ("max_y", C.c_float)]


# every symbol include/orb_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "orb_last_error", "orb_device_count", "orb_create", "orb_destroy", "orb_get_scale_tables",
    "orb_max_keypoints", "orb_max_keypoints_for_size", "orb_extract", "orb_extract_batch_host", "orb_extract_batch_host_async", "orb_extract_batch_device",
    "orb_extract_stereo", "orb_stereo_match", "orb_extract_stereo_batch_device",
    "orb_synchronize", "orb_last_launch_count", "orb_last_call_breakdown", "orb_set_profiling", "orb_get_stage_times", "orb_set_lanes", "orb_stage_level_size", "orb_stage_copy_level",
    "orb_stage_copy_blur", "orb_stage_copy_candidates", "orb_stage_copy_kept",
    "orb_compute_image_bounds", "orb_undistort_keypoints_device", "orb_assign_features_to_grid_device",
    "orb_get_features_in_area_device",
    "orb_descriptor_distance", "orb_matcher_create", "orb_matcher_destroy",
    "orb_search_for_initialization", "orb_match_pairs_device", "orb_match_allpairs_device",
    "orb_shard_range", "orb_nccl_unique_id", "orb_nccl_comm_create", "orb_nccl_comm_destroy", "orb_match_allpairs_nccl",
    "orb_hamming_matrix_device", "orb_matcher_synchronize", "orb_int_pipe_peak",
    "orb_search_scratch_bytes", "orb_project_last_frame_device", "orb_search_by_projection_device", "orb_search_by_projection_last_frame", "orb_search_by_projection_host",
    "orb_search_by_bow_device", "orb_search_for_triangulation_device", "orb_search_by_bow_host", "orb_search_for_triangulation_host",
    "orb_cvt_color_gray_device", "orb_remap_linear_device", "orb_distinctive_descriptors_device",
]

_lib = None


class OrbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("orb_b200 status %d: %s" % (status, msg))
        self.status = status


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("liborb_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` or `make -C orb_slam2_detailed_comments_b200/csrc`. There is no CPU "
                              "fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.orb_last_error.restype = C.c_char_p
        vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.orb_create.argtypes = [C.POINTER(OrbParams), i32, i32, C.POINTER(vp)]
        L.orb_destroy.argtypes = [vp]
        L.orb_get_scale_tables.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orb_max_keypoints.argtypes = [vp]
        L.orb_max_keypoints_for_size.argtypes = [vp, i32, i32]
        L.orb_extract.argtypes = [vp, vp, i32, i32, sz, vp, i32, C.POINTER(i32), vp, vp]
        L.orb_extract_batch_host.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, i32, vp, vp]
        L.orb_extract_batch_host_async.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, i32, vp, vp]
        L.orb_extract_batch_device.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, i32, vp, vp, vp]
        L.orb_extract_stereo.argtypes = [vp, vp, vp, i32, i32, sz, f32, f32, vp, i32, C.POINTER(i32), vp, vp, C.POINTER(i32), vp, vp, vp]
        L.orb_stereo_match.argtypes = [vp, vp, f32, f32, vp, vp, C.POINTER(i32)]
        L.orb_extract_stereo_batch_device.argtypes = [vp, vp, i32, i32, i32, sz, sz, vp, i32, vp, vp, f32, f32, vp, vp, vp]
        L.orb_synchronize.argtypes = [vp, vp]
        L.orb_last_launch_count.argtypes = [vp]
        L.orb_last_call_breakdown.argtypes = [vp, vp]
        L.orb_set_profiling.argtypes = [vp, i32]
        L.orb_set_lanes.argtypes = [vp, i32]
        L.orb_get_stage_times.argtypes = [vp, vp, vp]
        L.orb_stage_level_size.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.orb_stage_copy_level.argtypes = [vp, i32, i32, vp]
        L.orb_stage_copy_blur.argtypes = [vp, i32, i32, vp]
        L.orb_stage_copy_candidates.argtypes = [vp, i32, i32, vp, vp, vp, i32, C.POINTER(i32)]
        L.orb_stage_copy_kept.argtypes = [vp, i32, i32, vp, vp, vp, i32, C.POINTER(i32)]
        L.orb_compute_image_bounds.argtypes = [C.POINTER(OrbCamera), i32, i32, vp]
        L.orb_undistort_keypoints_device.argtypes = [i32, vp, vp, i32, i32, C.POINTER(OrbCamera), vp, vp]
        L.orb_assign_features_to_grid_device.argtypes = [i32, vp, vp, i32, i32, vp, vp, vp, vp]
        L.orb_get_features_in_area_device.argtypes = [i32, vp, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp]
        L.orb_descriptor_distance.argtypes = [vp, vp]
        L.orb_matcher_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
        L.orb_matcher_destroy.argtypes = [vp]
        L.orb_search_for_initialization.argtypes = [vp, C.POINTER(OrbFrameView), C.POINTER(OrbFrameView),
                                                    C.POINTER(OrbMatchParams), vp, vp, C.POINTER(i32), vp, vp]
        L.orb_match_pairs_device.argtypes = [vp, vp, vp, i32, i32, f32, i32, vp, vp, vp]
        L.orb_match_allpairs_device.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp]
        L.orb_shard_range.restype = None
        L.orb_shard_range.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.orb_nccl_unique_id.argtypes = [vp]
        L.orb_nccl_comm_create.argtypes = [i32, i32, i32, vp, C.POINTER(vp)]
        L.orb_nccl_comm_destroy.argtypes = [vp]
        L.orb_match_allpairs_nccl.argtypes = [vp, vp, i32, i32, vp, i32, i32, f32, vp, vp, vp]
        L.orb_hamming_matrix_device.argtypes = [vp, vp, i32, vp, i32, vp, vp]
        L.orb_matcher_synchronize.argtypes = [vp, vp]
        L.orb_int_pipe_peak.argtypes = [i32, i32, C.POINTER(C.c_double)]
        L.orb_search_for_triangulation_device.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, C.POINTER(OrbDeviceFrames), vp, vp, vp, vp, i32,
                                                          i32, vp, vp, vp, vp]
        L.orb_cvt_color_gray_device.argtypes = [i32, vp, i32, i32, sz, sz, i32, i32, vp, sz, sz, vp]
        L.orb_remap_linear_device.argtypes = [i32, vp, i32, i32, sz, sz, i32, vp, vp, i32, i32, vp, sz, sz, vp]
        L.orb_distinctive_descriptors_device.argtypes = [i32, vp, vp, i32, i32, vp, vp, vp]
        L.orb_search_scratch_bytes.restype = sz
        L.orb_search_scratch_bytes.argtypes = [i32, i32, i32]
        L.orb_project_last_frame_device.argtypes = [i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, f32, f32, vp, i32, vp, vp]
        L.orb_search_by_projection_device.argtypes = [i32, C.POINTER(OrbDeviceFrames), vp, vp, vp, i32, C.POINTER(OrbSearchParams),
                                                      vp, vp, vp, vp, vp]
        L.orb_search_by_bow_device.argtypes = [i32, vp, vp, vp, vp, vp, i32, C.POINTER(OrbDeviceFrames), vp,
                                               C.POINTER(OrbSearchParams), vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def check(status):
    if status != ORB_OK:
        raise OrbError(status, lib().orb_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Address of a numpy array or a torch tensor (device or host)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: names the legacy default stream explicitly (NULL means "the handle's own stream" in this ABI)


def stream_arg(stream, *tensors):
    """cudaStream_t to pass for a device entry point. An explicit `stream` (raw cudaStream_t int) wins. Otherwise, when
    the operands are torch CUDA tensors, the call is ordered on torch's CURRENT stream of their device - the stream the
    tensors were produced on (torch.zeros, copy_, NCCL work.wait()) - and not on the handle's private non-blocking
    stream, which is not ordered with it. Without torch tensors: NULL = the handle's own stream."""
    if stream:
        return C.c_void_p(int(stream))
    for t in tensors:
        if t is not None and not isinstance(t, np.ndarray) and getattr(t, "is_cuda", False):
            import torch
            s = torch.cuda.current_stream(t.device).cuda_stream
            return C.c_void_p(int(s) if s else CUDA_STREAM_LEGACY)
    return C.c_void_p(0)
