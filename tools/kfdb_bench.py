"""Wall-clock of one loop-candidate detection (KeyFrameDatabase::DetectLoopCandidates) through the drop-in on liborbgpu.so
against the reference's own KeyFrameDatabase.cc on one host core, for growing maps; prints one JSON line per size."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import kfdb_ref
import voc_ref

path = os.path.join(tempfile.mkdtemp(), "voc.txt")
voc_ref.write_text_vocabulary(path, k=10, L=5, seed=3)
for nkf in (500, 2000, 8000):
    out = {"keyframes": nkf}
    for kind in ("ref", "shim_gpu"):
        if not kfdb_ref.available(kind):
            continue
        lib = kfdb_ref.KfdbLib(kind, path)
        nwords = lib.L.voc_words(lib.voc)
        bows = kfdb_ref.bow_world(nwords, nkf, 5, places=60, per_kf=700, place_words=1200, noise_words=200)
        for k, (ids, w) in enumerate(bows):
            lib.new_keyframe(k + 1, ids, w)
        rng = np.random.default_rng(1)
        for k in range(nkf):
            lib.set_graph(k, *kfdb_ref.graph_of(nkf, k, rng))
        for k in range(nkf - 20):
            lib.add(k)
        ts, found = [], 0
        for rep in range(12):
            q = nkf - 1 - rep % 10
            t = time.perf_counter()
            r = lib.detect_loop(q, 0.01)
            ts.append(time.perf_counter() - t)
            found += len(r)
        out[kind + "_ms"] = float(np.median(ts[2:]) * 1e3)
        out[kind + "_candidates"] = found
        out["words_per_keyframe"] = float(np.mean([len(b[0]) for b in bows]))
        lib.close()
    print(json.dumps(out), flush=True)
