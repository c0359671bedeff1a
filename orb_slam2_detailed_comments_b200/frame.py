"""Host-side mirror of the Frame steps that follow extraction (src/Frame.cc of the reference):
ComputeImageBounds :779, UndistortKeyPoints :724, AssignFeaturesToGrid :399, GetFeaturesInArea :590.
The device entry points work on torch tensors that hold extraction output (capacity-strided)."""
import ctypes as C

import numpy as np

from ._lib import stream_arg, AREA_QUERY_DTYPE, OrbCamera, check, lib, ptr

FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


def camera(fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0):
    return OrbCamera(fx, fy, cx, cy, k1, k2, p1, p2, k3)


def ComputeImageBounds(cam, width, height):
    """-> np.float32[4] = (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
    b = np.zeros(4, np.float32)
    check(lib().orb_compute_image_bounds(C.byref(cam), width, height, ptr(b)))
    return b


def UndistortKeyPoints(d_kps, d_counts, cam, d_kps_un, device=0, stream=None):
    """d_kps / d_kps_un: (B, cap, 28) u8 torch tensors (orb_keypoint records), d_counts (B) int32."""
    B, cap = d_kps.shape[0], d_kps.shape[1]
    check(lib().orb_undistort_keypoints_device(device, ptr(d_kps), ptr(d_counts), B, cap, C.byref(cam), ptr(d_kps_un),
                                               stream_arg(stream, d_kps)))


def AssignFeaturesToGrid(d_kps_un, d_counts, bounds, d_cell_start, d_cell_items, device=0, stream=None):
    """d_cell_start (B, 3073) int32, d_cell_items (B, cap) int32: mGrid[ix][iy] = items[start[ix*48+iy] : start[ix*48+iy+1]]."""
    B, cap = d_kps_un.shape[0], d_kps_un.shape[1]
    b = np.ascontiguousarray(bounds, np.float32)
    check(lib().orb_assign_features_to_grid_device(device, ptr(d_kps_un), ptr(d_counts), B, cap, ptr(b), ptr(d_cell_start),
                                                   ptr(d_cell_items), stream_arg(stream, d_kps_un)))


def GetFeaturesInArea(d_kps_un, bounds, d_cell_start, d_cell_items, d_queries, d_out, d_out_counts, device=0, stream=None):
    """d_queries: (Q, 24) u8 tensor of AREA_QUERY_DTYPE records; d_out (Q, out_cap) int32; d_out_counts (Q) int32."""
    cap = d_kps_un.shape[1]
    b = np.ascontiguousarray(bounds, np.float32)
    check(lib().orb_get_features_in_area_device(device, ptr(d_kps_un), cap, ptr(b), ptr(d_cell_start), ptr(d_cell_items),
                                                ptr(d_queries), d_queries.shape[0], ptr(d_out), d_out.shape[1],
                                                ptr(d_out_counts), stream_arg(stream, d_kps_un)))


def make_queries(frames, xs, ys, rs, min_levels, max_levels):
    q = np.zeros(len(xs), AREA_QUERY_DTYPE)
    q["frame"], q["x"], q["y"], q["r"], q["min_level"], q["max_level"] = frames, xs, ys, rs, min_levels, max_levels
    return q
