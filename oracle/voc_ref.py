"""TEST INFRASTRUCTURE — synthetic DBoW2 vocabularies in the reference's text format (ORBvoc.txt, read by
ORBVocabulary::loadFromTextFile, src/ORBVocabulary.cc:11-139) and a ctypes view of oracle/_ref/lib{ref,shim}_voc*.so
(oracle/ref_shim/voc_api.cpp): the reference's own ORBVocabulary + DBoW2, alone or with shim/ORBVocabularyGPU.h."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_p, _i = C.c_void_p, C.c_int


def lib_path(kind):
    return os.path.join(HERE, "_ref", {"ref": "libref_voc.so", "shim_cpu": "libshim_voc_cpu.so", "shim_gpu": "libshim_voc_gpu.so"}[kind])


def available(kind):
    return os.path.exists(lib_path(kind))


def _ptr(a):
    return a.ctypes.data_as(_p)


def write_text_vocabulary(path, k=10, L=3, seed=0, shallow_leaf_fraction=0.05, stopped_fraction=0.03, flips=24):
    """A k-ary tree of depth L in ORBvoc.txt layout: header `k L scoring weighting` (L1_NORM = 0, TF_IDF = 0), then one line
    per node in creation order: parent id, is-leaf flag, 32 descriptor bytes, weight.  Children perturb their parent's
    descriptor (so the descent is meaningful), some inner-level nodes are made leaves (k-means stops early on small clusters)
    and some words get weight 0 ("stopped")."""
    rng = np.random.default_rng(seed)
    lines = []
    nodes = [(0, None, 0)]                        # (id, descriptor bits, level)
    frontier = [(0, rng.integers(0, 2, 256, dtype=np.uint8), 0)]
    next_id = 1
    while frontier:
        new = []
        for pid, bits, level in frontier:
            for _ in range(k):
                child = bits.copy()
                child[rng.choice(256, flips, replace=False)] ^= 1
                leaf = level + 1 == L or (level + 1 >= 2 and rng.random() < shallow_leaf_fraction)
                weight = 0.0 if (leaf and rng.random() < stopped_fraction) else float(rng.uniform(0.5, 9.0))
                desc = np.packbits(child)
                lines.append("%d %d %s %r" % (pid, 1 if leaf else 0, " ".join(str(int(b)) for b in desc), weight))
                if not leaf:
                    new.append((next_id, child, level + 1))
                next_id += 1
        frontier = new
    with open(path, "w") as f:
        f.write("%d %d 0 0\n" % (k, L))
        f.write("\n".join(lines) + "\n")
    return next_id


class VocLib:
    def __init__(self, kind, path):
        self.L = C.CDLL(lib_path(kind))
        L = self.L
        L.voc_load_text.restype = _p
        L.voc_load_text.argtypes = [C.c_char_p]
        L.voc_destroy.argtypes = [_p]
        for f in ("voc_nodes", "voc_words", "voc_depth"):
            getattr(L, f).argtypes = [_p]
        L.voc_export.argtypes = [_p] * 6
        L.voc_transform_each.argtypes = [_p, _p, _i, _i, _p, _p, _p]
        L.voc_transform.argtypes = [_p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p]
        self.h = L.voc_load_text(path.encode())
        assert self.h, "the reference's loader rejected %s" % path
        self.nodes, self.words, self.depth = L.voc_nodes(self.h), L.voc_words(self.h), L.voc_depth(self.h)

    def close(self):
        if self.h:
            self.L.voc_destroy(self.h)
            self.h = None

    def export(self):
        """Flat arrays of the tree as the reference's loader built it."""
        n = self.nodes
        start, kids = np.zeros(n + 1, np.int32), np.zeros(max(n - 1, 1), np.int32)
        desc, word, weight = np.zeros((n, 32), np.uint8), np.zeros(n, np.int32), np.zeros(n, np.float64)
        self.L.voc_export(self.h, _ptr(start), _ptr(kids), _ptr(desc), _ptr(word), _ptr(weight))
        return start, kids[:start[-1]], desc, word, weight

    def transform_each(self, desc, levelsup):
        d = np.ascontiguousarray(desc, np.uint8)
        n = len(d)
        words, weights, nids = np.zeros(n, np.int32), np.zeros(n, np.float64), np.zeros(n, np.uint32)
        self.L.voc_transform_each(self.h, _ptr(d), n, levelsup, _ptr(words), _ptr(weights), _ptr(nids))
        return words, weights, nids

    def transform(self, desc, levelsup):
        d = np.ascontiguousarray(desc, np.uint8)
        n = len(d)
        bow_ids, bow_vals = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
        fv_nodes, fv_off, fv_idx = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32)
        nbow, nfv = C.c_int(), C.c_int()
        self.L.voc_transform(self.h, _ptr(d), n, levelsup, _ptr(bow_ids), _ptr(bow_vals), C.byref(nbow), _ptr(fv_nodes), _ptr(fv_off), _ptr(fv_idx),
                             C.byref(nfv))
        return dict(bow_ids=bow_ids[:nbow.value].copy(), bow_vals=bow_vals[:nbow.value].copy(), fv_nodes=fv_nodes[:nfv.value].copy(),
                    fv_offsets=fv_off[:nfv.value + 1].copy(), fv_idx=fv_idx[:fv_off[nfv.value]].copy())


def features_near_words(desc, word, rng, n, flips=30):
    """Descriptors near randomly chosen leaves (plus pure noise), so that descents are not all decided at the root."""
    leaves = np.nonzero(word >= 0)[0]
    pick = leaves[rng.integers(0, len(leaves), n)]
    bits = np.unpackbits(desc[pick], axis=1)
    for i in range(n):
        bits[i, rng.choice(256, int(rng.integers(0, flips + 1)), replace=False)] ^= 1
    out = np.packbits(bits, axis=1)
    noise = rng.random(n) < 0.1
    out[noise] = rng.integers(0, 256, (int(noise.sum()), 32), dtype=np.uint8)
    return out
