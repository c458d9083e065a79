"""ctypes binding of liborbgpu.so (the C ABI in include/orbgpu.h).

The library is the product: there is no CPU fallback anywhere in this package.  If the shared
object has not been built (``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C refactored_orb_slam2_b200/csrc``) importing a compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORBGPU_LIB") or os.path.join(_HERE, "liborbgpu.so")    # ORBGPU_LIB: tuning builds (tools/)

OK, E_INVALID, E_CUDA, E_CAPACITY, E_UNSUPPORTED = 0, -1, -2, -3, -4

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])   # cv::KeyPoint, 28 bytes
TOP2_DTYPE = np.dtype([("d1", "<i4"), ("d2", "<i4"), ("idx", "<i8")])                # shard exchange record, 16 bytes
assert KP_DTYPE.itemsize == 28 and TOP2_DTYPE.itemsize == 16


class ExtractorParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("max_width", C.c_int32),
                ("max_height", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32)]


class FeaturesDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("u_right", C.c_void_p), ("descriptors", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("grid_w_inv", C.c_float), ("grid_h_inv", C.c_float),
                ("grid_cols", C.c_int32), ("grid_rows", C.c_int32), ("query_min_x", C.c_float), ("query_min_y", C.c_float),
                ("device", C.c_int32)]


class Camera(C.Structure):
    """orbgpu_camera: mK and mDistCoef of src/Frame.cc as the settings files give them (float)."""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 5), ("ndist", C.c_int32)]

    @classmethod
    def make(cls, fx, fy, cx, cy, dist):
        d = [float(np.float32(x)) for x in dist]
        if len(d) not in (4, 5):
            raise ValueError("4 or 5 distortion coefficients (k1 k2 p1 p2 [k3])")
        c = cls(float(np.float32(fx)), float(np.float32(fy)), float(np.float32(cx)), float(np.float32(cy)))
        for i, x in enumerate(d):
            c.dist[i] = x
        c.ndist = len(d)
        return c


class SearchPolicy(C.Structure):
    _fields_ = [("source", C.c_int32), ("ratio_mode", C.c_int32), ("greedy", C.c_int32), ("gate", C.c_int32),
                ("accept_max", C.c_int32), ("ratio", C.c_float), ("stereo_gate", C.c_int32), ("only_stereo", C.c_int32),
                ("ex", C.c_float), ("ey", C.c_float), ("f12", C.c_float * 9), ("level_sigma2", C.c_float * 16),
                ("inv_level_sigma2", C.c_float * 16), ("scale_factors", C.c_float * 16)]


QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"), ("ur", "<f4"),
                        ("ur_tol", "<f4"), ("list_begin", "<i4"), ("list_end", "<i4"), ("flags", "<i4")])     # orbgpu_query, 40 bytes
MATCH_DTYPE = np.dtype([("idx", "<i4"), ("dist", "<i4"), ("dist2", "<i4"), ("level", "i1"), ("level2", "i1"), ("accepted", "i1"),
                        ("pad", "i1")])                                                                        # orbgpu_match, 16 bytes
assert QUERY_DTYPE.itemsize == 40 and MATCH_DTYPE.itemsize == 16


class OrbGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orbgpu error %d: %s" % (code, msg))
        self.code = code


# every symbol include/orbgpu.h declares: name -> (restype, argtypes)
_vp, _i, _f, _sz, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64
SYMBOLS = {
    "orbgpu_version": (_i, []),
    "orbgpu_last_error": (C.c_char_p, []),
    "orbgpu_device_count": (_i, []),
    "orbgpu_extractor_create": (_i, [C.POINTER(ExtractorParams), C.POINTER(_vp)]),
    "orbgpu_extractor_destroy": (None, [_vp]),
    "orbgpu_extractor_tables": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "orbgpu_extractor_capacity": (_i, [_vp]),
    "orbgpu_extract": (_i, [_vp, _vp, _i, _i, _sz, _vp, _vp, _i, C.POINTER(_i)]),
    "orbgpu_extract_batch": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _vp, _vp, _i, _vp]),
    "orbgpu_extract_batch_async": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz, _vp, _vp, _i, _vp]),
    "orbgpu_wait": (_i, [_vp]),
    "orbgpu_extract_batch_device": (_i, [_vp, _vp, _i, _i, _i, _sz, _sz]),
    "orbgpu_results_device": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i)]),
    "orbgpu_download_results": (_i, [_vp, _vp, _vp, _i, _vp]),
    "orbgpu_synchronize": (_i, [_vp]),
    "orbgpu_extractor_stream": (_vp, [_vp]),
    "orbgpu_level_size": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "orbgpu_get_pyramid_level": (_i, [_vp, _i, _i, _i, _vp, _sz]),
    "orbgpu_get_stage": (_i, [_vp, _i, _i, _i, _vp, _i]),
    "orbgpu_get_blurred_level": (_i, [_vp, _i, _i, _vp, _sz]),
    "orbgpu_stereo_match": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _i]),
    "orbgpu_stereo_match_async": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _i]),
    "orbgpu_stereo_match_device": (_i, [_vp, _vp, _i, _i, _i, _f, _f]),
    "orbgpu_stereo_results_device": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "orbgpu_descriptor_distance": (_i, [_vp, _vp, _i, _vp, _i]),
    "orbgpu_bank_create": (_i, [_vp, _i, _i64, _i, C.POINTER(_vp)]),
    "orbgpu_bank_wrap_device": (_i, [_vp, _i, _i64, _i, C.POINTER(_vp)]),
    "orbgpu_bank_destroy": (None, [_vp]),
    "orbgpu_hamming_top2_dense": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "orbgpu_hamming_top2_dense_device": (_i, [_vp, _vp, _i, _vp, _vp]),
    "orbgpu_top2_merge_device": (_i, [_vp, _i, _i, _vp, _f, _i, _vp, _i, _vp]),
    "orbgpu_exchange_create": (_i, [_i, _i, _i, _i, C.POINTER(_vp), _vp]),
    "orbgpu_exchange_connect": (_i, [_vp, _vp]),
    "orbgpu_exchange_check": (_i, [_vp]),
    "orbgpu_exchange_destroy": (None, [_vp]),
    "orbgpu_hamming_top2_sharded_device": (_i, [_vp, _vp, _vp, _i, _vp, _f, _i, _vp, _vp]),
    "orbgpu_hamming_top2_csr": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "orbgpu_features_create": (_i, [C.POINTER(FeaturesDesc), C.POINTER(_vp)]),
    "orbgpu_features_from_extractor": (_i, [_vp, _i, _i, C.POINTER(FeaturesDesc), C.POINTER(_vp)]),
    "orbgpu_features_from_extractor_undistorted": (_i, [_vp, _i, _i, C.POINTER(FeaturesDesc), C.POINTER(Camera), C.POINTER(_vp)]),
    "orbgpu_undistort_keypoints": (_i, [_vp, _i, C.POINTER(Camera), _vp, _i]),
    "orbgpu_undistort_points": (_i, [_vp, _i, C.POINTER(Camera), _vp, _i]),
    "orbgpu_features_destroy": (None, [_vp]),
    "orbgpu_features_count": (_i, [_vp]),
    "orbgpu_features_grid": (_i, [_vp, _vp, _vp, C.POINTER(C.c_int32)]),
    "orbgpu_features_download": (_i, [_vp, _vp, _vp, _vp]),
    "orbgpu_search": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, C.POINTER(SearchPolicy), _vp, C.POINTER(_i)]),
    "orbgpu_features_in_area": (_i, [_vp, _vp, _i, _vp, _vp, _i]),
    "orbgpu_vocabulary_create": (_i, [C.c_int32, _vp, _vp, _vp, _vp, C.c_int32, _i, C.POINTER(_vp)]),
    "orbgpu_vocabulary_destroy": (None, [_vp]),
    "orbgpu_vocabulary_transform": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "orbgpu_vocabulary_transform_features": (_i, [_vp, _vp, _i, _vp, _vp]),
    "orbgpu_cvt_gray_device": (_i, [_vp, _i, _i, _i, _i, _sz, _sz, _vp, _sz, _sz, _i, _vp]),
    "orbgpu_remap_device": (_i, [_vp, _i, _i, _sz, _sz, _vp, _vp, _i, _i, _vp, _sz, _sz, _i, _vp]),
    "orbgpu_cvt_gray": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i]),
    "orbgpu_remap": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _i, _i]),
    "orbgpu_distinctive_descriptors": (_i, [_vp, _vp, _i, _vp, _vp, _i]),
    "orbgpu_bowdb_create": (_i, [_i, C.POINTER(_vp)]),
    "orbgpu_bowdb_destroy": (None, [_vp]),
    "orbgpu_bowdb_add": (_i, [_vp, _vp, _vp, _i]),
    "orbgpu_bowdb_erase": (_i, [_vp, _i]),
    "orbgpu_bowdb_clear": (_i, [_vp]),
    "orbgpu_bowdb_slots": (_i, [_vp]),
    "orbgpu_bowdb_query": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i]),
    "orbgpu_selftest_sincos": (_i, [C.c_uint32, C.c_uint64, _i, _vp, _vp, _vp, _i]),
    "orbgpu_measure_popc_peak": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "orbgpu_measure_pipe": (_i, [_i, _i, C.POINTER(C.c_double)]),
    "orbgpu_launch_count": (_i64, [_i]),
    "orbgpu_set_profiling": (_i, [_vp, _i]),
    "orbgpu_stage_times": (_i, [_vp, _vp, _i]),
    "orbgpu_stage_name": (C.c_char_p, [_i]),
}

_LIB = None


def lib():
    """Loads liborbgpu.so once; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `make -C %s/csrc` (or __graft_entry__.build()); "
                              "this package has no CPU fallback" % (LIB_PATH, _HERE))
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)      # AttributeError if the ABI and the header drifted apart
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc < 0:
        raise OrbGpuError(rc, lib().orbgpu_last_error().decode("utf-8", "replace"))
    return rc


def ptr(a):
    """void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
