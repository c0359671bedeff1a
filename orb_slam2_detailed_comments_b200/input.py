"""Host-side mirror of the input stage in front of the extractor and of the map-point descriptor choice:
cv::cvtColor to gray (src/Tracking.cc:250-276), cv::remap rectification (Examples/Stereo/stereo_euroc.cc:181-188),
MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:365-448). torch CUDA tensors in, torch CUDA tensors out."""
import ctypes as C

from ._lib import stream_arg, check, lib, ptr

RGB2GRAY, BGR2GRAY, RGBA2GRAY, BGRA2GRAY = 0, 1, 2, 3


def cvtColorGray(d_src, code, d_gray, device=0, stream=None):
    """d_src (B, H, W, 3|4) u8, d_gray (B, H, W) u8."""
    B, H, W, ch = d_src.shape
    assert ch == (4 if code in (RGBA2GRAY, BGRA2GRAY) else 3)
    assert d_src.stride(3) == 1 and d_src.stride(2) == ch and d_gray.stride(2) == 1, "pixels must be dense inside a row"
    check(lib().orb_cvt_color_gray_device(device, ptr(d_src), W, H, d_src.stride(1), d_src.stride(0), B, code, ptr(d_gray),
                                          d_gray.stride(1), d_gray.stride(0), stream_arg(stream, d_src)))


def remap(d_src, d_map_x, d_map_y, d_dst, device=0, stream=None):
    """cv::remap(src, dst, mapx, mapy, INTER_LINEAR): d_src (B, Hs, Ws) u8, maps (Hd, Wd) f32, d_dst (B, Hd, Wd) u8."""
    B, Hs, Ws = d_src.shape
    Hd, Wd = d_map_x.shape
    assert d_src.stride(2) == 1 and d_dst.stride(2) == 1 and d_map_x.is_contiguous() and d_map_y.is_contiguous()
    check(lib().orb_remap_linear_device(device, ptr(d_src), Ws, Hs, d_src.stride(1), d_src.stride(0), B, ptr(d_map_x), ptr(d_map_y),
                                        Wd, Hd, ptr(d_dst), d_dst.stride(1), d_dst.stride(0), stream_arg(stream, d_src)))


def ComputeDistinctiveDescriptors(d_desc, d_offsets, max_observations, d_best_index, d_best_desc=None, device=0, stream=None):
    """d_desc (total, 32) u8, d_offsets (P+1) int32 -> d_best_index (P) int32 [, d_best_desc (P, 32) u8]."""
    check(lib().orb_distinctive_descriptors_device(device, ptr(d_desc), ptr(d_offsets), d_offsets.shape[0] - 1, int(max_observations),
                                                   ptr(d_best_index), ptr(d_best_desc), stream_arg(stream, d_desc)))
