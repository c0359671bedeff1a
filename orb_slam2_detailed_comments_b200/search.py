"""Host-side mirror of the ordered candidate-set matchers of the reference's ORBmatcher:
SearchByProjection (local map, src/ORBmatcher.cc:72-169; last frame, :1710-1860) and
SearchByBoW(KeyFrame*, Frame&) (:247-420). All tensors are torch CUDA tensors, capacity-strided
per frame; the kernels live in csrc/orb_search.cu. There is no CPU path."""
import ctypes as C

import numpy as np

from ._lib import (stream_arg, ORB_SEARCH_BEST, ORB_SEARCH_RATIO, ORB_SEARCH_RATIO_LEVEL, PROJ_QUERY_DTYPE, TRI_PAIR_DTYPE, OrbDeviceFrames,
                   OrbSearchParams, check, lib, ptr)

TH_HIGH, TH_LOW = 100, 50   # src/ORBmatcher.cc:47-49


def scratch_bytes(batch, query_capacity, capacity):
    return int(lib().orb_search_scratch_bytes(batch, query_capacity, capacity))


def device_frames(d_kps_un, d_desc, d_counts, bounds, d_cell_start=None, d_cell_items=None, d_uright=None, d_occupied=None):
    """Describes `batch` current frames; the returned struct keeps the tensors alive."""
    f = OrbDeviceFrames()
    f.keypoints_un = ptr(d_kps_un); f.descriptors = ptr(d_desc); f.counts = ptr(d_counts)
    f.uright = ptr(d_uright); f.occupied = ptr(d_occupied)
    f.cell_start = ptr(d_cell_start); f.cell_items = ptr(d_cell_items)
    b = np.ascontiguousarray(bounds, np.float32)
    for i in range(4):
        f.bounds[i] = float(b[i])
    f.batch, f.capacity = d_kps_un.shape[0], d_kps_un.shape[1]
    f._keep = (d_kps_un, d_desc, d_counts, d_uright, d_occupied, d_cell_start, d_cell_items)
    return f


def _mul3(a0, b0, a1, b1, a2, b2):
    f = np.float32
    return f(f(f(a0 * b0) + f(a1 * b1)) + f(a2 * b2))


def motion_direction(Tcw_cur, Tcw_last, mb, mono):
    """bForward / bBackward of ORBmatcher.cc:1717-1730 -> 1 / 2, else 0 (float32, cv::gemm order)."""
    if mono:
        return 0
    Tc = np.asarray(Tcw_cur, np.float32); Tl = np.asarray(Tcw_last, np.float32)
    Rcw, tcw, Rlw, tlw = Tc[:3, :3], Tc[:3, 3], Tl[:3, :3], Tl[:3, 3]
    twc = [np.float32(-_mul3(Rcw[0, r], tcw[0], Rcw[1, r], tcw[1], Rcw[2, r], tcw[2])) for r in range(3)]
    tlc_z = np.float32(_mul3(Rlw[2, 0], twc[0], Rlw[2, 1], twc[1], Rlw[2, 2], twc[2]) + tlw[2])
    if tlc_z > np.float32(mb):
        return 1
    if -tlc_z > np.float32(mb):
        return 2
    return 0


def local_map_queries(proj_x, proj_y, proj_xr, level, view_cos, in_view, nobs_positive, th, scale_factors):
    """Queries of SearchByProjection(F, vpMapPoints, th) from the MapPoint track fields Frame::isInFrustum wrote
    (mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos, mbTrackInView && !isBad()); r as ORBmatcher.cc:88-100."""
    f = np.float32
    level = np.asarray(level, np.int32)
    r = np.where(np.asarray(view_cos, f) > f(0.998), f(2.5), f(4.0)).astype(f)   # RadiusByViewingCos :171-177
    if th != 1.0:
        r = (r * f(th)).astype(f)
    q = np.zeros(len(level), PROJ_QUERY_DTYPE)
    q["u"], q["v"], q["ur"] = proj_x, proj_y, proj_xr
    q["radius"] = (r * np.asarray(scale_factors, f)[level]).astype(f)
    q["min_level"], q["max_level"] = level - 1, level
    q["flags"] = np.asarray(in_view, np.int32) | (np.asarray(nobs_positive, np.int32) << 1)
    return q


def ProjectLastFrame(d_world_pos, d_mp_flags, d_last_kps, d_last_counts, d_Tcw, d_direction, cam4, bounds, mbf, th,
                     scale_factors, d_queries, device=0, stream=None):
    B, qcap = d_last_kps.shape[0], d_last_kps.shape[1]
    cam = np.ascontiguousarray(cam4, np.float32); b = np.ascontiguousarray(bounds, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    check(lib().orb_project_last_frame_device(device, ptr(d_world_pos), ptr(d_mp_flags), ptr(d_last_kps), ptr(d_last_counts), B, qcap,
                                              ptr(d_Tcw), ptr(d_direction), ptr(cam), ptr(b), float(mbf), float(th), ptr(sf),
                                              len(sf), ptr(d_queries), stream_arg(stream, d_world_pos)))


def _params(mode, th, ratio, check_ori):
    return OrbSearchParams(int(mode), int(th), float(ratio), int(bool(check_ori)))


def SearchByProjection(frames, d_queries, d_query_desc, d_query_counts, mode, th, nn_ratio, check_orientation, d_scratch,
                       d_match_of_keypoint, d_match_of_query, d_nmatches, device=0, stream=None):
    """mode ORB_SEARCH_RATIO_LEVEL + TH_HIGH = local map (:72); ORB_SEARCH_BEST + TH_HIGH + orientation = last frame (:1710)."""
    p = _params(mode, th, nn_ratio, check_orientation)
    check(lib().orb_search_by_projection_device(device, C.byref(frames), ptr(d_queries), ptr(d_query_desc), ptr(d_query_counts),
                                                d_queries.shape[1], C.byref(p), ptr(d_scratch), ptr(d_match_of_keypoint),
                                                ptr(d_match_of_query), ptr(d_nmatches), stream_arg(stream, d_queries)))


def SearchByBoW(d_kps1, d_desc1, d_node1, d_usable1, d_counts1, frames, d_node2, nn_ratio, check_orientation, d_scratch,
                d_match_of_keypoint, d_match_of_query, d_nmatches, th=TH_LOW, device=0, stream=None):
    p = _params(ORB_SEARCH_RATIO, th, nn_ratio, check_orientation)
    check(lib().orb_search_by_bow_device(device, ptr(d_kps1), ptr(d_desc1), ptr(d_node1), ptr(d_usable1), ptr(d_counts1),
                                         d_kps1.shape[1], C.byref(frames), ptr(d_node2), C.byref(p), ptr(d_scratch),
                                         ptr(d_match_of_keypoint), ptr(d_match_of_query), ptr(d_nmatches),
                                         stream_arg(stream, d_kps1)))


def epipole(K2_fx, K2_fy, K2_cx, K2_cy, R2w, t2w, Cw):
    """ex, ey of SearchForTriangulation (ORBmatcher.cc:893-901): keyframe 1's camera centre Cw in keyframe 2's image
    (float32, cv::gemm order, invz = 1.0f / z in float)."""
    f = np.float32
    R, t, C_ = np.asarray(R2w, f), np.asarray(t2w, f).reshape(3), np.asarray(Cw, f).reshape(3)
    c2 = [f(_mul3(R[r, 0], C_[0], R[r, 1], C_[1], R[r, 2], C_[2]) + t[r]) for r in range(3)]
    invz = f(f(1.0) / c2[2])
    return f(f(f(f(K2_fx) * c2[0]) * invz) + f(K2_cx)), f(f(f(f(K2_fy) * c2[1]) * invz) + f(K2_cy))


def SearchForTriangulation(d_kps1, d_desc1, d_node1, d_has_mp1, d_uright1, d_counts1, frames2, d_node2, d_pairs, scale_factors,
                           level_sigma2, check_orientation, d_scratch, d_matches12, d_nmatches, device=0, stream=None):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:884) for a batch of keyframe pairs. frames2 = device_frames(mvKeysUn,
    mDescriptors, N, bounds (unused), uright = mvuRight, occupied = "has a map point") of keyframe 2; d_pairs: (B, 48) u8 tensor of
    TRI_PAIR_DTYPE records."""
    sf = np.ascontiguousarray(scale_factors, np.float32); s2 = np.ascontiguousarray(level_sigma2, np.float32)
    check(lib().orb_search_for_triangulation_device(device, ptr(d_kps1), ptr(d_desc1), ptr(d_node1), ptr(d_has_mp1), ptr(d_uright1),
                                                    ptr(d_counts1), d_kps1.shape[1], C.byref(frames2), ptr(d_node2), ptr(d_pairs), ptr(sf),
                                                    ptr(s2), len(sf), int(bool(check_orientation)), ptr(d_scratch), ptr(d_matches12),
                                                    ptr(d_nmatches), stream_arg(stream, d_kps1)))
