"""ctypes front-end to oracle/_ref/libref_orb.so — the reference's OWN, unmodified ORBextractor.cc
compiled (from /root/reference, by oracle/ref_shim/Makefile) against an OpenCV API stand-in whose image
primitives are the cv2-pinned restatements of liborb_oracle.so, with a monotone allocator so that its
pointer sort is deterministic.  TEST INFRASTRUCTURE ONLY.  The .so is built in this container only
(the GPU box has no /root/reference) and travels with the snapshot."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_orb.so")
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def available():
    """True if the checker exists or can be built here."""
    if os.path.exists(SO):
        return True
    return os.path.isdir("/root/reference/Source/Libraries/ORB_SLAM2")


def lib():
    global _LIB
    if _LIB is None:
        if os.path.isdir("/root/reference/Source/Libraries/ORB_SLAM2"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref_shim")])
        if not os.path.exists(SO):
            raise ImportError("oracle/_ref/libref_orb.so is missing and /root/reference is not here to build it")
        C.CDLL(os.path.join(_HERE, "liborb_oracle.so"), mode=C.RTLD_GLOBAL)
        L = C.CDLL(SO)
        L.ref_orb_create.restype = C.c_void_p
        L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_orb_destroy.argtypes = [C.c_void_p]
        L.ref_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_orb_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_orb_stereo_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_long, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


class ReferenceExtractor:
    """ORB_SLAM2::ORBextractor itself (ORBextractor.h:47-57)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = int(nfeatures), int(nlevels)
        self._h = C.c_void_p(lib().ref_orb_create(int(nfeatures), float(scale_factor), int(nlevels), int(ini_th), int(min_th)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_orb_destroy(self._h)
            self._h = None

    def __call__(self, image, mask=None):
        img = np.ascontiguousarray(image, np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 3 * self.nlevels + 256
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().ref_orb_extract(self._h, img.ctypes.data_as(C.c_void_p), w, h, img.strides[0],
                                  kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap)
        if n < 0:
            raise RuntimeError("capacity too small: need %d" % -n)
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        lib().ref_orb_level(self._h, l, None, 0, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        lib().ref_orb_level(self._h, l, out.ctypes.data_as(C.c_void_p), out.strides[0], C.byref(w), C.byref(h))
        return out


def stereo_batch(params, left, right, mbf, mb, nthreads):
    """One stereo pair per thread: the reference's extraction L + R, then the ComputeStereoMatches restatement."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    n, h, w = left.shape
    cap = params[0] + 3 * params[2] + 256
    cl, cr, m = (np.zeros(n, np.int32) for _ in range(3))
    lib().ref_orb_stereo_batch(int(params[0]), float(params[1]), int(params[2]), int(params[3]), int(params[4]),
                               left.ctypes.data_as(C.c_void_p), right.ctypes.data_as(C.c_void_p), n, w, h, h * w, float(mbf),
                               float(mb), cap, cl.ctypes.data_as(C.c_void_p), cr.ctypes.data_as(C.c_void_p),
                               m.ctypes.data_as(C.c_void_p), int(nthreads))
    return cl, cr, m
