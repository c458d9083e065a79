"""ctypes front-end to oracle/liborb_oracle.so (the C++ CPU restatement of the reference path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)


def build(force=False):
    so = os.path.join(_HERE, "liborb_oracle.so")
    srcs = [os.path.join(_HERE, "oracle_cpu.cpp"), os.path.join(_HERE, "oracle_search.cpp")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liborb_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orb_oracle_create.restype = C.c_void_p
        L.orb_oracle_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orb_oracle_destroy.argtypes = [C.c_void_p]
        L.orb_oracle_set_tie_rule.argtypes = [C.c_void_p, C.c_int]
        L.orb_oracle_tables.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _i32p, _i32p]
        L.orb_oracle_extract.restype = C.c_int
        L.orb_oracle_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_int]
        L.orb_oracle_extract_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.c_void_p,
                                               C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_level_size.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
        L.orb_oracle_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_get_stage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orb_oracle_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_fast_atan2.restype = C.c_float
        L.orb_oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orb_oracle_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_sincos.argtypes = [C.c_float, _f32p, _f32p]
        L.orb_oracle_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p, C.c_int]
        L.orb_oracle_hamming.argtypes = [C.c_void_p, C.c_void_p]
        L.orb_oracle_top2_dense.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int]
        L.orb_oracle_top2_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.orb_oracle_accept.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
        L.orb_oracle_rotation_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orb_oracle_stereo.restype = C.c_int
        L.orb_oracle_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        L.orb_oracle_stereo_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, C.c_int, C.c_long, C.c_float, C.c_float, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orb_oracle_features_create.restype = C.c_void_p
        L.orb_oracle_features_create.argtypes = [C.c_void_p]
        L.orb_oracle_features_destroy.argtypes = [C.c_void_p]
        L.orb_oracle_features_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orb_oracle_features_in_area.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orb_oracle_voc_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orb_oracle_cvt_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orb_oracle_remap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orb_oracle_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orb_oracle_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleExtractor:
    """Mirror of ORB_SLAM2::ORBextractor (ORBextractor.h:47-57) on the CPU oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.params = (int(nfeatures), float(scale_factor), int(nlevels), int(ini_th), int(min_th))
        self.nfeatures, self.nlevels = int(nfeatures), int(nlevels)
        self._h = C.c_void_p(lib().orb_oracle_create(*self.params))
        sc, isc, s2, is2 = (np.zeros(nlevels, np.float32) for _ in range(4))
        q, um = np.zeros(nlevels, np.int32), np.zeros(16, np.int32)
        lib().orb_oracle_tables(self._h, sc.ctypes.data_as(_f32p), isc.ctypes.data_as(_f32p),
                                s2.ctypes.data_as(_f32p), is2.ctypes.data_as(_f32p),
                                q.ctypes.data_as(_i32p), um.ctypes.data_as(_i32p))
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2, self.quota, self.umax = sc, isc, s2, is2, q, um

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orb_oracle_destroy(self._h)
            self._h = None

    def set_tie_rule(self, lo):
        lib().orb_oracle_set_tie_rule(self._h, int(lo))

    @property
    def capacity(self):
        # a level can return 4 x (octree roots) keypoints even when its quota is smaller (wide images, small nfeatures)
        return self.nfeatures + 3 * self.nlevels + 64 * self.nlevels

    def __call__(self, image, mask=None):
        img = np.ascontiguousarray(image, np.uint8)
        h, w = img.shape
        cap = self.capacity
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().orb_oracle_extract(self._h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap)
        if n < 0:
            raise RuntimeError("oracle capacity too small: need %d" % -n)
        return kps[:n].copy(), desc[:n].copy()

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        lib().orb_oracle_level_size(self._h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, level, which=0):
        """which: 0 pyramid interior, 1 with the 19 px border, 2 blurred interior."""
        w, h = self.level_size(level)
        b = 19 if which == 1 else 0
        out = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
        rc = lib().orb_oracle_get_level(self._h, level, which, _p(out), out.strides[0])
        return out if rc == 0 else None

    def stage(self, level, stage):
        """stage 0: FAST candidates in reference order; 1: octree selection in list order.
        Rows (x, y, response) in octree coordinates (level coordinates - 16)."""
        n = lib().orb_oracle_get_stage(self._h, level, stage, None, 0)
        out = np.zeros((max(n, 1), 3), np.int32)
        lib().orb_oracle_get_stage(self._h, level, stage, _p(out), n)
        return out[:n]

    def distribute(self, cands, min_x, max_x, min_y, max_y, n_want):
        cands = np.ascontiguousarray(cands, np.int32).reshape(-1, 3)
        out = np.zeros((len(cands) + 4, 3), np.int32)
        n = lib().orb_oracle_distribute(self._h, _p(cands), len(cands), min_x, max_x, min_y, max_y, n_want,
                                        _p(out), len(out))
        return out[:n]


def extract_batch(params, imgs, nthreads):
    """One frame per thread (Frame.cc:87-90 threading model). imgs: [n, h, w] u8."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    cap = params[0] + 3 * params[2] + 64
    kps = np.zeros((n, cap), KP_DTYPE)
    desc = np.zeros((n, cap, 32), np.uint8)
    counts = np.zeros(n, np.int32)
    lib().orb_oracle_extract_batch(*params, _p(imgs), n, w, h, w, h * w, _p(kps), _p(desc), cap, _p(counts), nthreads)
    return kps, desc, counts


def stereo_batch(params, left, right, mbf, mb, nthreads):
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    n, h, w = left.shape
    cap = params[0] + 3 * params[2] + 64
    cl, cr, m = (np.zeros(n, np.int32) for _ in range(3))
    lib().orb_oracle_stereo_batch(*params, _p(left), _p(right), n, w, h, h * w, mbf, mb, cap, _p(cl), _p(cr), _p(m),
                                  nthreads)
    return cl, cr, m


def stereo(ext_l, ext_r, kl, dl, kr, dr, mbf, mb):
    """Frame::ComputeStereoMatches (Frame.cc:477-646) on the pyramids of the two oracles' last calls."""
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl); dr = np.ascontiguousarray(dr)
    n = len(kl)
    u = np.zeros(n, np.float32); z = np.zeros(n, np.float32); bd = np.zeros(n, np.int32)
    kept = lib().orb_oracle_stereo(ext_l._h, ext_r._h, _p(kl), _p(dl), n, _p(kr), _p(dr), len(kr), mbf, mb,
                                   _p(u), _p(z), _p(bd))
    return u, z, bd, kept


def undistort_points(xy, fx, fy, cx, cy, dist):
    """cv::undistortPoints(xy, K, dist, R=None, P=K) restated (Frame.cc:419-445); xy float32 [n, 2]."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    d = np.ascontiguousarray(dist, np.float32).ravel()
    out = np.empty_like(xy)
    L = lib()
    L.orb_oracle_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
    L.orb_oracle_undistort_points.restype = None
    L.orb_oracle_undistort_points(_p(xy), len(xy), fx, fy, cx, cy, _p(d), len(d), _p(out))
    return out


def bow_query(database, query, alive=None):
    """KeyFrameDatabase candidate scoring for every keyframe of `database` (list of (ascending word ids, weights)) against
    `query`: (shared words, first shared word, DBoW2 L1 score as float64)."""
    n = len(database)
    start = np.zeros(n + 1, np.int32)
    start[1:] = np.cumsum([len(i) for i, _ in database])
    words = np.concatenate([np.asarray(i, np.uint32) for i, _ in database]) if n else np.zeros(0, np.uint32)
    weights = np.concatenate([np.asarray(w, np.float64) for _, w in database]) if n else np.zeros(0)
    live = np.ones(n, np.uint8) if alive is None else np.ascontiguousarray(alive, np.uint8)
    qi, qw = np.ascontiguousarray(query[0], np.uint32), np.ascontiguousarray(query[1], np.float64)
    common, first, score = np.zeros(n, np.int32), np.zeros(n, np.uint32), np.zeros(n, np.float64)
    L = lib()
    L.orb_oracle_bow_query.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3
    L.orb_oracle_bow_query.restype = None
    L.orb_oracle_bow_query(n, _p(start), _p(np.ascontiguousarray(words)), _p(np.ascontiguousarray(weights)), _p(live), _p(qi), _p(qw), len(qi),
                           _p(common), _p(first), _p(score))
    return common, first, score


def hamming(a, b):
    return lib().orb_oracle_hamming(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)))


def top2_dense(q, bank, nthreads=1):
    q = np.ascontiguousarray(q, np.uint8); bank = np.ascontiguousarray(bank, np.uint8)
    nq = len(q)
    i1, d1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
    lib().orb_oracle_top2_dense(_p(q), nq, _p(bank), len(bank), _p(i1), _p(d1), _p(d2), nthreads)
    return i1, d1, d2


def top2_csr(q, train, offsets, cand):
    q = np.ascontiguousarray(q, np.uint8); train = np.ascontiguousarray(train, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32); cand = np.ascontiguousarray(cand, np.int32)
    nq = len(q)
    i1, d1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
    lib().orb_oracle_top2_csr(_p(q), nq, _p(train), _p(offsets), _p(cand), _p(i1), _p(d1), _p(d2))
    return i1, d1, d2


def accept(d1, d2, nnratio=0.7, th_low=50):
    return np.array([lib().orb_oracle_accept(int(a), int(b), nnratio, th_low) for a, b in zip(d1, d2)], bool)


def rotation_filter(angle1, angle2):
    a1 = np.ascontiguousarray(angle1, np.float32); a2 = np.ascontiguousarray(angle2, np.float32)
    keep = np.zeros(len(a1), np.uint8)
    lib().orb_oracle_rotation_filter(_p(a1), _p(a2), len(a1), _p(keep))
    return keep.astype(bool)


def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orb_oracle_resize(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def blur(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orb_oracle_blur(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast(img, th):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 2 + 16
    out = np.zeros((cap, 3), np.int32)
    n = lib().orb_oracle_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], th, _p(out), cap)
    return out[:n]


def fast_atan2(y, x):
    return lib().orb_oracle_fast_atan2(float(y), float(x))


def sincos(angle):
    c, s = C.c_float(), C.c_float()
    lib().orb_oracle_sincos(float(angle), C.byref(c), C.byref(s))
    return c.value, s.value


class OracleFeatures:
    """CPU restatement (oracle_search.cpp) of a frame's grid and of the ORBmatcher candidate loops; same record
    layouts as include/orbgpu.h (the ctypes structures are shared with the product's binding)."""

    def __init__(self, keys_un, descriptors, u_right, geometry):
        from refactored_orb_slam2_b200._lib import KP_DTYPE, FeaturesDesc
        k = np.ascontiguousarray(keys_un, KP_DTYPE)
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        u = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        desc = FeaturesDesc()
        C.memmove(C.byref(desc), C.byref(geometry), C.sizeof(FeaturesDesc))
        desc.n, desc.keys_un, desc.descriptors = len(k), _p(k), _p(d)
        desc.u_right = _p(u) if u is not None else None
        self.n = len(k)
        self.ncells = desc.grid_cols * desc.grid_rows
        self._h = C.c_void_p(lib().orb_oracle_features_create(C.byref(desc)))

    def __del__(self):
        try:
            lib().orb_oracle_features_destroy(self._h)
        except Exception:
            pass

    def grid(self):
        start = np.empty(self.ncells + 1, np.int32)
        items = np.empty(max(self.n, 1), np.int32)
        n = C.c_int32()
        lib().orb_oracle_features_grid(self._h, _p(start), _p(items), C.byref(n))
        return start, items[:n.value].copy()

    def in_area(self, queries):
        q = np.ascontiguousarray(queries)
        offsets = np.empty(len(q) + 1, np.int32)
        items = np.empty(max(1, len(q) * max(self.n, 1)), np.int32) if len(q) * self.n < (1 << 24) else np.empty(1 << 24, np.int32)
        lib().orb_oracle_features_in_area(self._h, _p(q), len(q), _p(offsets), _p(items), len(items))
        return offsets, items[:offsets[-1]].copy()

    def search(self, queries, query_descriptors, policy, occupied=None, cand=None):
        from refactored_orb_slam2_b200._lib import MATCH_DTYPE
        q = np.ascontiguousarray(queries)
        d = np.ascontiguousarray(query_descriptors, np.uint8).reshape(-1, 32)
        occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        cd = np.zeros(1, np.int32) if cand is None else np.ascontiguousarray(cand, np.int32)
        out = np.zeros(len(q), MATCH_DTYPE)
        rc = lib().orb_oracle_search(self._h, _p(q), _p(d), len(q), _p(occ) if occ is not None else None, _p(cd), 0 if cand is None else len(cd),
                                     C.byref(policy), _p(out))
        assert rc == 0, rc
        return out


def voc_transform(child_start, children, node_desc, word_id, depth_levels, features, levelsup):
    """Restated DBoW2 tree descent (oracle_search.cpp): (word_ids, node_ids) per feature."""
    cs, ch = np.ascontiguousarray(child_start, np.int32), np.ascontiguousarray(children, np.int32)
    nd, wi = np.ascontiguousarray(node_desc, np.uint8), np.ascontiguousarray(word_id, np.int32)
    f = np.ascontiguousarray(features, np.uint8).reshape(-1, 32)
    words, nodes = np.zeros(len(f), np.int32), np.zeros(len(f), np.int32)
    rc = lib().orb_oracle_voc_transform(len(wi), _p(cs), _p(ch), _p(nd), _p(wi), int(depth_levels), _p(f), len(f), int(levelsup), _p(words), _p(nodes))
    assert rc == 0
    return words, nodes


def cvt_gray(img, code):
    """Restated cvtColor(BGR/RGB[A] -> GRAY) for one image [H, W, 3|4]; code 0 = BGR order, 1 = RGB order."""
    a = np.ascontiguousarray(img, np.uint8)
    out = np.empty(a.shape[:2], np.uint8)
    lib().orb_oracle_cvt_gray(_p(a), a.shape[2], int(code), a.shape[1], a.shape[0], _p(out))
    return out


def remap(img, map_x, map_y):
    """Restated cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) with float32 maps."""
    a = np.ascontiguousarray(img, np.uint8)
    mx, my = np.ascontiguousarray(map_x, np.float32), np.ascontiguousarray(map_y, np.float32)
    out = np.empty(mx.shape, np.uint8)
    lib().orb_oracle_remap(_p(a), a.shape[1], a.shape[0], _p(mx), _p(my), mx.shape[1], mx.shape[0], _p(out))
    return out


def distinctive(descriptors, offsets):
    """Restated MapPoint::ComputeDistinctiveDescriptors over groups of rows: (best index per group, its median)."""
    d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(offsets, np.int32)
    g = len(off) - 1
    idx, med = np.zeros(g, np.int32), np.zeros(g, np.int32)
    lib().orb_oracle_distinctive(_p(d) if len(d) else None, _p(off), g, _p(idx), _p(med))
    return idx, med
