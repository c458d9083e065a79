"""Latency of ONE KITTI-shaped stereo pair through the host-buffer C ABI (BASELINE.json configs[1] as the reference runs it:
a frame at a time), next to the CPU reference on one pair (two threads, as Frame.cc:87-90)."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch

import gen_synth
import refactored_orb_slam2_b200 as pkg

W, H = 1241, 376
PARAMS = (2000, 1.2, 8, 20, 7)
BF, FX = 386.1448, 718.856
left, right = gen_synth.frame_batch(8, H, W, 100, stereo=True)
lib = pkg.lib()
vp = ctypes.c_void_p
mb = float(np.float32(np.float32(BF) / np.float32(FX)))
ex = pkg.ORBextractor(*PARAMS, max_width=W, max_height=H, max_batch=2)
cap = ex.capacity
host = torch.empty((2, H, W), dtype=torch.uint8).pin_memory()
k = torch.empty((2, cap, 28), dtype=torch.uint8).pin_memory()
d = torch.empty((2, cap, 32), dtype=torch.uint8).pin_memory()
c = torch.empty(2, dtype=torch.int32).pin_memory()
u = torch.empty((1, cap), dtype=torch.float32).pin_memory()
z = torch.empty((1, cap), dtype=torch.float32).pin_memory()


def one(i):
    host[0].copy_(torch.from_numpy(left[i % 8])); host[1].copy_(torch.from_numpy(right[i % 8]))
    t = time.perf_counter()
    pkg._lib.check(lib.orbgpu_extract_batch(ex.handle, vp(host.data_ptr()), 2, W, H, W, W * H, vp(k.data_ptr()), vp(d.data_ptr()), cap, vp(c.data_ptr())))
    pkg._lib.check(lib.orbgpu_stereo_match(ex.handle, ex.handle, 0, 1, 1, BF, mb, vp(u.data_ptr()), vp(z.data_ptr()), cap))
    return time.perf_counter() - t


for i in range(10):
    one(i)
ts = sorted(one(i) for i in range(200))
lib.orbgpu_launch_count(1)
one(0)
launches = lib.orbgpu_launch_count(0)
out = {"what": "one KITTI stereo pair: 2 x extraction + ComputeStereoMatches, host buffers in / out", "median_ms": round(ts[100] * 1e3, 3), "p90_ms": round(ts[180] * 1e3, 3),
       "min_ms": round(ts[0] * 1e3, 3), "kernel_launches": int(launches), "keypoints": [int(c[0]), int(c[1])]}
try:
    import ref
    if ref.available():
        l8, r8 = left[:1], right[:1]
        ref.stereo_batch(PARAMS, l8, r8, BF, np.float32(mb), 1)
        t = time.perf_counter()
        for _ in range(5):
            ref.stereo_batch(PARAMS, l8, r8, BF, np.float32(mb), 1)
        out["cpu_reference_ms_one_thread"] = round((time.perf_counter() - t) / 5 * 1e3, 2)
except Exception as e:
    out["cpu_reference_error"] = str(e)
print(json.dumps(out))
