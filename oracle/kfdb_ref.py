"""TEST INFRASTRUCTURE — ctypes view of oracle/_ref/lib{ref,shim}_kfdb*.so (oracle/ref_shim/kfdb_api.cpp): the reference's
own KeyFrameDatabase.cc, or the drop-in shim/KeyFrameDatabase.cc, over the reference's real vocabulary code, plus a synthetic
keyframe world (places with characteristic words, a covisibility graph along the trajectory)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_p, _i = C.c_void_p, C.c_int


def lib_path(kind):
    return os.path.join(HERE, "_ref", {"ref": "libref_kfdb.so", "shim_cpu": "libshim_kfdb_cpu.so", "shim_gpu": "libshim_kfdb_gpu.so"}[kind])


def available(kind):
    return os.path.exists(lib_path(kind))


def _ptr(a):
    return a.ctypes.data_as(_p)


class KfdbLib:
    def __init__(self, kind, voc_path):
        if kind == "shim_cpu":
            C.CDLL(os.path.join(HERE, "liborb_oracle.so"), mode=C.RTLD_GLOBAL)
        self.L = L = C.CDLL(lib_path(kind))
        L.voc_load_text.restype = _p
        L.voc_load_text.argtypes = [C.c_char_p]
        L.voc_destroy.argtypes = [_p]
        L.kfdb_create.restype = _p
        L.kfdb_create.argtypes = [_p]
        L.kfdb_destroy.argtypes = [_p]
        L.kfdb_new_keyframe.argtypes = [_p, C.c_ulong, _p, _p, _i]
        L.kfdb_set_graph.argtypes = [_p, _i, _p, _i, _p, _i]
        for f in ("kfdb_add", "kfdb_erase"):
            getattr(L, f).argtypes = [_p, _i]
        L.kfdb_clear.argtypes = [_p]
        L.kfdb_detect_loop.argtypes = [_p, _i, C.c_float, _p, _i]
        L.kfdb_detect_reloc.argtypes = [_p, C.c_ulong, _p, _p, _i, _p, _i]
        L.kfdb_state.argtypes = [_p, _i, _p, _p]
        L.kfdb_voc_score.restype = C.c_double
        L.kfdb_voc_score.argtypes = [_p, _p, _p, _i, _p, _p, _i]
        self.voc = L.voc_load_text(voc_path.encode())
        assert self.voc
        self.h = L.kfdb_create(self.voc)
        self.n = 0

    def close(self):
        if self.h:
            self.L.kfdb_destroy(self.h)
            self.L.voc_destroy(self.voc)
            self.h = None

    def new_keyframe(self, kf_id, ids, weights):
        ids, weights = np.ascontiguousarray(ids, np.uint32), np.ascontiguousarray(weights, np.float64)
        self.n += 1
        return self.L.kfdb_new_keyframe(self.h, int(kf_id), _ptr(ids), _ptr(weights), len(ids))

    def set_graph(self, kf, connected, covisible):
        c, v = np.ascontiguousarray(connected, np.int32), np.ascontiguousarray(covisible, np.int32)
        self.L.kfdb_set_graph(self.h, kf, _ptr(c), len(c), _ptr(v), len(v))

    def add(self, kf):
        self.L.kfdb_add(self.h, kf)

    def erase(self, kf):
        self.L.kfdb_erase(self.h, kf)

    def clear(self):
        self.L.kfdb_clear(self.h)

    def detect_loop(self, kf, min_score):
        out = np.full(self.n + 1, -9, np.int32)
        n = self.L.kfdb_detect_loop(self.h, kf, float(min_score), _ptr(out), len(out))
        return out[:n].copy()

    def detect_reloc(self, frame_id, ids, weights):
        ids, weights = np.ascontiguousarray(ids, np.uint32), np.ascontiguousarray(weights, np.float64)
        out = np.full(self.n + 1, -9, np.int32)
        n = self.L.kfdb_detect_reloc(self.h, int(frame_id), _ptr(ids), _ptr(weights), len(ids), _ptr(out), len(out))
        return out[:n].copy()

    def state(self):
        """[n, 4] int64 (loopQuery, loopWords, relocQuery, relocWords) and [n, 2] float32 (loopScore, relocScore)."""
        ints, sc = np.zeros((self.n, 4), np.int64), np.zeros((self.n, 2), np.float32)
        for k in range(self.n):
            self.L.kfdb_state(self.h, k, _ptr(ints[k]), _ptr(sc[k]))
        return ints, sc

    def voc_score(self, a, b):
        (i1, w1), (i2, w2) = a, b
        i1, w1 = np.ascontiguousarray(i1, np.uint32), np.ascontiguousarray(w1, np.float64)
        i2, w2 = np.ascontiguousarray(i2, np.uint32), np.ascontiguousarray(w2, np.float64)
        return self.L.kfdb_voc_score(self.h, _ptr(i1), _ptr(w1), len(i1), _ptr(i2), _ptr(w2), len(i2))


def bow_world(nwords, nkf, seed, places=9, place_words=350, per_kf=220, noise_words=60):
    """BowVectors of `nkf` keyframes along a trajectory that visits `places` places twice (so loops exist): ascending word ids
    and L1-normalised positive weights, like a tf-idf BowVector after DBoW2's normalisation."""
    rng = np.random.default_rng(seed)
    vocab = [np.sort(rng.choice(nwords, place_words, replace=False)) for _ in range(places)]
    out = []
    for k in range(nkf):
        p = (k * 2 * places // nkf) % places
        ids = np.union1d(rng.choice(vocab[p], per_kf, replace=False), rng.choice(nwords, noise_words, replace=False)).astype(np.uint32)
        w = rng.uniform(0.2, 6.0, len(ids)) * rng.integers(1, 4, len(ids))
        out.append((ids, w / w.sum()))
    return out


def graph_of(nkf, k, rng):
    connected = [j for j in range(max(0, k - 4), min(nkf, k + 5)) if j != k]
    covisible = sorted(connected, key=lambda j: (abs(j - k), j))
    extra = [int(j) for j in rng.choice(nkf, 3, replace=False) if j != k and j not in connected]
    return connected, (covisible + extra)[:12]
