"""Latency of the DBoW2 tree descent for one frame (2000 features, k = 10, L = 6-like tree) through the C ABI, next to the CPU oracle.
Usage: python tools/voc_bench.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import oracle
import refactored_orb_slam2_b200 as pkg
import test_gpu_vocabulary as T
import voc_ref

for k, L in ((10, 4), (10, 5)):
    start, kids, desc, word = T.synthetic_tree(k, L, seed=1, shallow=0.03)
    voc = pkg.Vocabulary(start, kids, desc, word, L)
    for n in (2000, 8000):
        feats = voc_ref.features_near_words(desc, word, np.random.default_rng(0), n)
        for _ in range(3):
            gw, gn = voc.transform(feats, 4)
        t = time.perf_counter()
        for _ in range(20):
            gw, gn = voc.transform(feats, 4)
        gpu = (time.perf_counter() - t) / 20
        t = time.perf_counter()
        ow, on = oracle.voc_transform(start, kids, desc, word, L, feats, 4)
        cpu = time.perf_counter() - t
        print(json.dumps({"k": k, "L": L, "nodes": int(len(desc)), "features": n, "hamming_per_feature": k * L, "gpu_call_us": round(gpu * 1e6, 1),
                          "cpu_oracle_us": round(cpu * 1e6, 1), "equal": bool(np.array_equal(gw, ow) and np.array_equal(gn, on))}))
