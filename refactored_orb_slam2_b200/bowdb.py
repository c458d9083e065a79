"""BowDatabase — the BowVectors of a KeyFrameDatabase resident in HBM and the candidate scoring of
KeyFrameDatabase::DetectLoopCandidates / DetectRelocalizationCandidates (Source/Libraries/ORB_SLAM2/src/KeyFrameDatabase.cc:72-306):
one query returns, for every keyframe, the shared-word count, the first shared word and DBoW2's L1 score.  The C++ drop-in
(shim/KeyFrameDatabase.cc) is the production host side; this is the same boundary for Python callers and tests."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr


class BowDatabase:
    def __init__(self, device=-1):
        self._h = C.c_void_p()
        check(lib().orbgpu_bowdb_create(int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().orbgpu_bowdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().orbgpu_bowdb_slots(self._h)

    def add(self, word_ids, weights):
        """KeyFrameDatabase::add: ascending word ids + tf-idf weights of one keyframe's BowVector; returns its slot."""
        w = np.ascontiguousarray(word_ids, np.uint32)
        v = np.ascontiguousarray(weights, np.float64)
        if len(w) != len(v):
            raise ValueError("word ids and weights differ in length")
        return check(lib().orbgpu_bowdb_add(self._h, ptr(w), ptr(v), len(w)))

    def erase(self, slot):
        check(lib().orbgpu_bowdb_erase(self._h, int(slot)))

    def clear(self):
        check(lib().orbgpu_bowdb_clear(self._h))

    def query(self, word_ids, weights):
        """(common words [slots] int32, first common word [slots] uint32, L1 score [slots] float64)."""
        w = np.ascontiguousarray(word_ids, np.uint32)
        v = np.ascontiguousarray(weights, np.float64)
        n = len(self)
        common, first, score = np.zeros(n, np.int32), np.zeros(n, np.uint32), np.zeros(n, np.float64)
        check(lib().orbgpu_bowdb_query(self._h, ptr(w), ptr(v), len(w), ptr(common), ptr(first), ptr(score), n))
        return common, first, score
