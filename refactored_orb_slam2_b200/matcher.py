"""ORBmatcher — host-side mirror of the Hamming half of ORB_SLAM2::ORBmatcher.

The reference's twelve searches (Source/Libraries/ORB_SLAM2/src/ORBmatcher.cc) share one inner
shape: for each query descriptor scan a candidate list, keep best / second-best Hamming distance
with strict '<', then apply threshold / ratio / orientation policy.  The scans run in liborbgpu.so
(dense bank scan for BASELINE config 4, candidate-list scan for the projection / BoW searches);
the policy below follows ORBmatcher.cc:224-226 (acceptance), :232-239 (rotation histogram) and
:1506-1538 (ComputeThreeMaxima) and stays on the host like the pointer-graph code it belongs to.
"""
import ctypes as C

import numpy as np

from ._lib import TOP2_DTYPE, check, lib, ptr


class DescriptorBank:
    """Keyframe descriptors resident in HBM ([rows, 32] uint8).  `index_base` offsets the reported
    row numbers so a row-sharded bank returns global indices."""

    def __init__(self, descriptors=None, index_base=0, device=-1, *, device_ptr=None, rows=None, keepalive=None):
        out = C.c_void_p()
        if device_ptr is not None:
            self.rows = int(rows)
            check(lib().orbgpu_bank_wrap_device(C.c_void_p(int(device_ptr)), self.rows, int(index_base), device, C.byref(out)))
            self._keepalive = keepalive
        else:
            d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
            self.rows = len(d)
            check(lib().orbgpu_bank_create(ptr(d), self.rows, int(index_base), device, C.byref(out)))
        self._h = out
        self.index_base = int(index_base)
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().orbgpu_bank_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def top2(self, queries):
        """Best / second-best over the whole bank: (idx1 int64, d1, d2), idx1 = -1 for an empty bank."""
        q = np.ascontiguousarray(queries, np.uint8).reshape(-1, 32)
        n = len(q)
        idx1 = np.full(n, -1, np.int64)
        d1 = np.full(n, 256, np.int32)
        d2 = np.full(n, 256, np.int32)
        check(lib().orbgpu_hamming_top2_dense(self._h, ptr(q), n, ptr(idx1), ptr(d1), ptr(d2)))
        return idx1, d1, d2

    def top2_device(self, d_queries, nq, d_result, stream=None):
        """Device form: writes nq 16-byte {d1, d2, idx} records at d_result (asynchronous)."""
        check(lib().orbgpu_hamming_top2_dense_device(self._h, C.c_void_p(int(d_queries)), int(nq),
                                                     C.c_void_p(int(d_result)), C.c_void_p(stream or 0)))


def merge_top2_device(d_parts, parts, nq, d_result, nnratio=0.0, th_low=0, d_match=None, device=-1, stream=None):
    """Folds [parts][nq] shard records (ascending bank order) on the device; optional acceptance test."""
    check(lib().orbgpu_top2_merge_device(C.c_void_p(int(d_parts)), int(parts), int(nq),
                                         C.c_void_p(int(d_result)) if d_result else None, float(nnratio), int(th_low),
                                         C.c_void_p(int(d_match)) if d_match else None, device, C.c_void_p(stream or 0)))


def distinctive_descriptors(descriptors, offsets, device=-1):
    """MapPoint::ComputeDistinctiveDescriptors (Source/Libraries/ORB_SLAM2/src/MapPoint.cc:229-321) for many MapPoints at
    once: group g = rows offsets[g]:offsets[g+1] (observation order); returns (best row per group, relative; its median)."""
    d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(offsets, np.int32)
    g = len(off) - 1
    idx, med = np.zeros(g, np.int32), np.zeros(g, np.int32)
    check(lib().orbgpu_distinctive_descriptors(ptr(d) if len(d) else None, ptr(off), g, ptr(idx), ptr(med), device))
    return idx, med


class ORBmatcher:
    TH_HIGH = 100        # ORBmatcher.cc:38-40
    TH_LOW = 50
    HISTO_LENGTH = 30

    def __init__(self, nnratio=0.6, checkOri=True, device=-1):
        self.mfNNratio = np.float32(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.device = device

    @staticmethod
    def DescriptorDistance(a, b, device=-1):
        """ORBmatcher::DescriptorDistance (ORBmatcher.cc:1542-1556) for one pair or for row-aligned batches."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        if a.shape != b.shape:
            raise ValueError("descriptor batches must have equal shapes")
        out = np.zeros(len(a), np.int32)
        check(lib().orbgpu_descriptor_distance(ptr(a), ptr(b), len(a), ptr(out), device))
        return int(out[0]) if len(out) == 1 else out

    def accept(self, d1, d2, th=None):
        """ORBmatcher.cc:224-226: bestDist1 <= TH_LOW and bestDist1 < mfNNratio * bestDist2."""
        th = self.TH_LOW if th is None else th
        d1 = np.asarray(d1)
        return (d1 <= th) & (d1.astype(np.float32) < self.mfNNratio * np.asarray(d2).astype(np.float32))

    def SearchBank(self, queries, bank, th=None):
        """The SearchByBoW(KF, Frame) inner loop (ORBmatcher.cc:201-243) over ONE giant node: every
        query scans the whole bank.  Returns (match, d1, d2): match[i] = bank row or -1."""
        idx1, d1, d2 = bank.top2(queries)
        ok = self.accept(d1, d2, th)
        return np.where(ok, idx1, -1), d1, d2

    def SearchCandidates(self, queries, train, offsets, cand, th=None):
        """Candidate-list form (projection / BoW searches): query i scans train rows
        cand[offsets[i]:offsets[i+1]] in list order.  Returns (match, d1, d2)."""
        q = np.ascontiguousarray(queries, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        off = np.ascontiguousarray(offsets, np.int32)
        cd = np.ascontiguousarray(cand, np.int32)
        n = len(q)
        idx1, d1, d2 = (np.zeros(n, np.int32) for _ in range(3))
        check(lib().orbgpu_hamming_top2_csr(ptr(q), n, ptr(t), len(t), ptr(off), ptr(cd), ptr(idx1), ptr(d1), ptr(d2),
                                            self.device))
        ok = self.accept(d1, d2, th)
        return np.where(ok, idx1, -1), d1, d2

    # -- orientation consistency, ORBmatcher.cc:232-239, 255-270, 1506-1538 -------------------
    @classmethod
    def ComputeThreeMaxima(cls, histo):
        max1 = max2 = max3 = 0
        ind1 = ind2 = ind3 = -1
        for i, s in enumerate(histo):
            if s > max1:
                max3, max2, max1 = max2, max1, s
                ind3, ind2, ind1 = ind2, ind1, i
            elif s > max2:
                max3, max2 = max2, s
                ind3, ind2 = ind2, i
            elif s > max3:
                max3, ind3 = s, i
        if max2 < np.float32(0.1) * np.float32(max1):
            ind2 = ind3 = -1
        elif max3 < np.float32(0.1) * np.float32(max1):
            ind3 = -1
        return ind1, ind2, ind3

    @classmethod
    def rotation_consistent(cls, angle1, angle2):
        """keep[i] for matches with keypoint angles angle1[i], angle2[i] (degrees)."""
        rot = np.asarray(angle1, np.float32) - np.asarray(angle2, np.float32)
        rot = np.where(rot < 0.0, rot + np.float32(360.0), rot).astype(np.float32)
        factor = np.float32(1.0) / np.float32(cls.HISTO_LENGTH)
        v = (rot * factor).astype(np.float32)
        bins = np.where(v >= 0, np.floor(v + np.float32(0.5)), -np.floor(-v + np.float32(0.5))).astype(np.int64)   # C round()
        bins[bins == cls.HISTO_LENGTH] = 0
        hist = np.bincount(bins, minlength=cls.HISTO_LENGTH)
        keepers = set(cls.ComputeThreeMaxima(hist.tolist()))
        return np.array([b in keepers for b in bins], bool)
