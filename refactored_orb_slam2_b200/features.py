"""FrameFeatures — one Frame's / KeyFrame's keypoints, right coordinates, descriptors and 64x48 grid resident in HBM,
and `search`, the candidate loops of the reference's ORBmatcher routines run on it.

Mirrors Frame::AssignFeaturesToGrid / PosInGrid (Source/Libraries/ORB_SLAM2/src/Frame.cc:250-263, 399-411),
Frame::GetFeaturesInArea (:341-397), KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:526-567) and the loops of
src/ORBmatcher.cc listed in include/orbgpu.h.  The C++ drop-in (shim/ORBmatcher.cc) is the production host side; this
module is the same boundary for Python callers and tests.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, MATCH_DTYPE, QUERY_DTYPE, Camera, FeaturesDesc, SearchPolicy, check, lib, ptr

FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48          # include/Frame.h:35-36
SRC_WINDOW, SRC_LIST = 0, 1
RATIO_NONE, RATIO_SAME_LEVEL, RATIO_LESS = 0, 1, 2
GREEDY_NONE, GREEDY_TAKEN, GREEDY_DISTANCE = 0, 1, 2
GATE_NONE, GATE_FUSE, GATE_EPIPOLAR = 0, 1, 2
Q_BLOCKS, Q_STEREO = 1, 2


def grid_geometry(min_x, max_x, min_y, max_y, keyframe=False, device=-1):
    """Geometry block of a feature set for an image with undistorted bounds [min_x, max_x] x [min_y, max_y]
    (Frame.cc:108-111); keyframe=True makes the window walk read the truncated integer bounds of KeyFrame.h:190-193."""
    d = FeaturesDesc()
    d.min_x, d.min_y = min_x, min_y
    d.grid_w_inv = np.float32(FRAME_GRID_COLS) / (np.float32(max_x) - np.float32(min_x))
    d.grid_h_inv = np.float32(FRAME_GRID_ROWS) / (np.float32(max_y) - np.float32(min_y))
    d.grid_cols, d.grid_rows = FRAME_GRID_COLS, FRAME_GRID_ROWS
    d.query_min_x, d.query_min_y = (float(int(min_x)), float(int(min_y))) if keyframe else (min_x, min_y)
    d.device = device
    return d


def make_policy(source=SRC_WINDOW, ratio_mode=RATIO_NONE, greedy=GREEDY_NONE, accept_max=100, ratio=0.0, stereo_gate=0, gate=GATE_NONE,
                only_stereo=0, ex=0.0, ey=0.0, f12=None, level_sigma2=None, inv_level_sigma2=None, scale_factors=None):
    P = SearchPolicy()
    P.source, P.ratio_mode, P.greedy, P.gate, P.accept_max, P.ratio = source, ratio_mode, greedy, gate, accept_max, ratio
    P.stereo_gate, P.only_stereo, P.ex, P.ey = stereo_gate, only_stereo, ex, ey
    for name, v in (("f12", f12), ("level_sigma2", level_sigma2), ("inv_level_sigma2", inv_level_sigma2), ("scale_factors", scale_factors)):
        if v is not None:
            arr = getattr(P, name)
            for i, x in enumerate(np.asarray(v, np.float32).ravel()[:len(arr)]):
                arr[i] = float(x)
    return P


def UndistortKeyPoints(keys, camera, device=-1):
    """Frame::UndistortKeyPoints (Frame.cc:419-445): mvKeysUn = mvKeys with pt = cv::undistortPoints(pt, mK, mDistCoef, P = mK),
    a copy when mDistCoef[0] == 0.  `camera` is a _lib.Camera (Camera.make(fx, fy, cx, cy, [k1, k2, p1, p2, k3]))."""
    k = np.ascontiguousarray(keys, KP_DTYPE)
    out = np.empty_like(k)
    check(lib().orbgpu_undistort_keypoints(ptr(k), len(k), C.byref(camera), ptr(out), device))
    return out


def undistort_points(xy, camera, device=-1):
    """cv::undistortPoints on bare float32 (x, y) pairs — the image corners of Frame::ComputeImageBounds (Frame.cc:447-475)."""
    p = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty_like(p)
    check(lib().orbgpu_undistort_points(ptr(p), len(p), C.byref(camera), ptr(out), device))
    return out


def ComputeImageBounds(width, height, camera, device=-1):
    """Frame::ComputeImageBounds (Frame.cc:447-475): (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
    if camera is None or camera.dist[0] == 0.0:
        return np.float32(0), np.float32(width), np.float32(0), np.float32(height)
    m = undistort_points([[0, 0], [width, 0], [0, height], [width, height]], camera, device)
    return min(m[0, 0], m[2, 0]), max(m[1, 0], m[3, 0]), min(m[0, 1], m[1, 1]), max(m[2, 1], m[3, 1])


class FrameFeatures:
    def __init__(self, keys_un=None, descriptors=None, u_right=None, geometry=None, *, _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            k = np.ascontiguousarray(keys_un, KP_DTYPE)
            d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
            if len(k) != len(d):
                raise ValueError("keypoints and descriptors differ in length")
            u = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
            desc = FeaturesDesc()
            C.memmove(C.byref(desc), C.byref(geometry), C.sizeof(FeaturesDesc))
            desc.n, desc.keys_un, desc.descriptors, desc.u_right = len(k), ptr(k), ptr(d), ptr(u)
            out = C.c_void_p()
            check(lib().orbgpu_features_create(C.byref(desc), C.byref(out)))
            self._h = out
        self.n = lib().orbgpu_features_count(self._h)

    @classmethod
    def from_extractor(cls, extractor, image, geometry, pair=-1, camera=None):
        """Device-to-device from the last extract_batch of `extractor` (and row `pair` of its last stereo match).  With a
        `camera` whose k1 != 0 the keypoints are undistorted on the device first (Frame::UndistortKeyPoints)."""
        out = C.c_void_p()
        if camera is None:
            check(lib().orbgpu_features_from_extractor(extractor._h, int(image), int(pair), C.byref(geometry), C.byref(out)))
        else:
            check(lib().orbgpu_features_from_extractor_undistorted(extractor._h, int(image), int(pair), C.byref(geometry), C.byref(camera),
                                                                   C.byref(out)))
        return cls(_handle=out)

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().orbgpu_features_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def grid(self):
        """(cell_start [64*48+1], cell_items): cell (ix, iy) is row ix*48+iy, items in ascending keypoint index."""
        start = np.empty(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, np.int32)
        items = np.empty(max(self.n, 1), np.int32)
        n = C.c_int32()
        check(lib().orbgpu_features_grid(self._h, ptr(start), ptr(items), C.byref(n)))
        return start, items[:n.value].copy()

    def download(self):
        k, u, d = np.empty(self.n, KP_DTYPE), np.empty(self.n, np.float32), np.empty((self.n, 32), np.uint8)
        check(lib().orbgpu_features_download(self._h, ptr(k), ptr(u), ptr(d)))
        return k, u, d

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        """Batched Frame::GetFeaturesInArea: scalars or arrays; returns (offsets, items)."""
        q = np.zeros(np.broadcast(x, y, r, minLevel, maxLevel).shape or (1,), QUERY_DTYPE).ravel()
        q["u"], q["v"], q["r"], q["min_level"], q["max_level"] = x, y, r, minLevel, maxLevel
        offsets = np.empty(len(q) + 1, np.int32)
        cap = max(1024, 64 * len(q))
        while True:
            items = np.empty(cap, np.int32)
            rc = lib().orbgpu_features_in_area(self._h, ptr(q), len(q), ptr(offsets), ptr(items), cap)
            if rc == -3:
                cap = int(offsets[-1]) + 16
                continue
            check(rc)
            return offsets, items[:offsets[-1]].copy()

    def search(self, queries, query_descriptors, policy, occupied=None, cand=None):
        """Runs one routine's candidate loops; returns (matches as MATCH_DTYPE array, passes used)."""
        q = np.ascontiguousarray(queries, QUERY_DTYPE)
        d = np.ascontiguousarray(query_descriptors, np.uint8).reshape(-1, 32)
        if len(q) != len(d):
            raise ValueError("queries and descriptors differ in length")
        occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        if occ is not None and len(occ) != self.n:
            raise ValueError("occupied must have one byte per keypoint")
        cd = np.zeros(0, np.int32) if cand is None else np.ascontiguousarray(cand, np.int32)
        out = np.zeros(len(q), MATCH_DTYPE)
        it = C.c_int()
        check(lib().orbgpu_search(self._h, ptr(q), ptr(d), len(q), ptr(occ), ptr(cd), len(cd), C.byref(policy), ptr(out), C.byref(it)))
        return out, it.value
