"""Vocabulary — a DBoW2 vocabulary tree resident in HBM and the per-feature descent of
TemplatedVocabulary::transform (Source/ThirdParty/DBoW2/DBoW2-local/include/DBoW2/TemplatedVocabulary.h:1216-1258),
the work behind Frame::ComputeBoW (Source/Libraries/ORB_SLAM2/src/Frame.cc:412-417).  The C++ drop-in is
shim/ORBVocabularyGPU.h; this is the same boundary for Python callers and tests.  No CPU fallback."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr


class Vocabulary:
    def __init__(self, child_start, children, node_descriptors, word_id, depth_levels, device=-1):
        cs = np.ascontiguousarray(child_start, np.int32)
        ch = np.ascontiguousarray(children, np.int32)
        nd = np.ascontiguousarray(node_descriptors, np.uint8).reshape(-1, 32)
        wi = np.ascontiguousarray(word_id, np.int32)
        if len(cs) != len(nd) + 1 or len(wi) != len(nd):
            raise ValueError("child_start needs nodes + 1 entries, word_id and descriptors one per node")
        out = C.c_void_p()
        check(lib().orbgpu_vocabulary_create(len(nd), ptr(cs), ptr(ch) if len(ch) else None, ptr(nd), ptr(wi), int(depth_levels), device, C.byref(out)))
        self._h = out
        self.nodes, self.depth_levels = len(nd), int(depth_levels)

    def close(self):
        if getattr(self, "_h", None) is not None:
            lib().orbgpu_vocabulary_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, descriptors, levelsup=4):
        """(word_ids, node_ids) for host descriptors [n, 32] or for a FrameFeatures set already in HBM."""
        from .features import FrameFeatures
        if isinstance(descriptors, FrameFeatures):
            n = descriptors.n
            words, nodes = np.zeros(n, np.int32), np.zeros(n, np.int32)
            check(lib().orbgpu_vocabulary_transform_features(self._h, descriptors._h, int(levelsup), ptr(words), ptr(nodes)))
            return words, nodes
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        words, nodes = np.zeros(len(d), np.int32), np.zeros(len(d), np.int32)
        check(lib().orbgpu_vocabulary_transform(self._h, ptr(d), len(d), int(levelsup), ptr(words), ptr(nodes)))
        return words, nodes
