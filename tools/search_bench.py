"""Latency of one ORBmatcher search through the C ABI (host buffers in, matches out) next to the CPU oracle's loops.
Usage: python tools/search_bench.py  (prints one JSON line per policy)."""
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
import slam_ref
import test_gpu_search as T

W, V = T.make_view(7, npoints=2200, clutter=600)
geo = pkg.grid_geometry(0.0, 1241.0, 0.0, 376.0)
g, o = pkg.FrameFeatures(V.kp, V.desc, V.uR, geo), oracle.OracleFeatures(V.kp, V.desc, V.uR, geo)
rng = np.random.default_rng(1)
for nq in (500, 2000, 8000):
    q, desc, _ = T.random_window_queries(pkg, rng, V, nq, radius=(5, 40))
    for name, kw in (("local_points(greedy)", dict(ratio_mode=1, greedy=1, accept_max=100, ratio=0.8, stereo_gate=1)), ("sim3(independent)", dict(accept_max=100))):
        P = pkg.make_policy(**kw)
        for _ in range(3):
            got, passes = g.search(q, desc, P)
        t = time.perf_counter()
        for _ in range(20):
            got, passes = g.search(q, desc, P)
        gpu = (time.perf_counter() - t) / 20
        t = time.perf_counter()
        for _ in range(3):
            want = o.search(q, desc, P)
        cpu = (time.perf_counter() - t) / 3
        off, items = o.in_area(q)
        print(json.dumps({"policy": name, "keypoints": int(len(V.kp)), "queries": nq, "candidates": int(off[-1]), "passes": passes, "gpu_call_us": round(gpu * 1e6, 1),
                          "cpu_oracle_us": round(cpu * 1e6, 1), "equal": bool(np.array_equal(got, want))}))
