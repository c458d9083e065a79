"""Input staging on the device: cvtColor(*2GRAY) and remap(INTER_LINEAR) as the reference applies them before the extractor
(Source/Libraries/ORB_SLAM2/src/Tracking.cc:166-182; Source/Examples/Stereo/stereo_euroc.cc:109-111, 159-160).  Host-array
forms for tests and small jobs, device-pointer forms for pipelines that keep frames in HBM.  No CPU fallback."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

BGR2GRAY, RGB2GRAY = 0, 1


def cvtColorGray(images, code=BGR2GRAY, device=-1):
    """images uint8 [B, H, W, 3|4] or [H, W, 3|4] -> gray uint8 [B, H, W] / [H, W]."""
    a = np.ascontiguousarray(images, np.uint8)
    single = a.ndim == 3
    if single:
        a = a[None]
    out = np.empty(a.shape[:3], np.uint8)
    check(lib().orbgpu_cvt_gray(ptr(a), a.shape[3], int(code), a.shape[2], a.shape[1], ptr(out), a.shape[0], device))
    return out[0] if single else out


def remap(images, map_x, map_y, device=-1):
    """images uint8 [B, H, W] or [H, W]; float32 maps [h, w] -> uint8 [B, h, w] / [h, w]."""
    a = np.ascontiguousarray(images, np.uint8)
    single = a.ndim == 2
    if single:
        a = a[None]
    mx, my = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
    if mx.shape != my.shape or mx.ndim != 2:
        raise ValueError("map_x and map_y must be equal-shaped 2-D float32 arrays")
    out = np.empty((a.shape[0],) + mx.shape, np.uint8)
    check(lib().orbgpu_remap(ptr(a), a.shape[2], a.shape[1], ptr(mx), ptr(my), mx.shape[1], mx.shape[0], ptr(out), a.shape[0], device))
    return out[0] if single else out


def cvtColorGray_device(d_src, channels, code, width, height, d_dst, batch=1, src_step=None, dst_step=None, stream=None):
    src_step = src_step or width * channels
    dst_step = dst_step or width
    check(lib().orbgpu_cvt_gray_device(C.c_void_p(int(d_src)), channels, int(code), width, height, src_step, src_step * height, C.c_void_p(int(d_dst)),
                                       dst_step, dst_step * height, batch, C.c_void_p(stream or 0)))


def remap_device(d_src, src_width, src_height, d_map_x, d_map_y, width, height, d_dst, batch=1, stream=None):
    check(lib().orbgpu_remap_device(C.c_void_p(int(d_src)), src_width, src_height, src_width, src_width * src_height, C.c_void_p(int(d_map_x)),
                                    C.c_void_p(int(d_map_y)), width, height, C.c_void_p(int(d_dst)), width, width * height, batch, C.c_void_p(stream or 0)))
