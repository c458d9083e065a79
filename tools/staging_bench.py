"""Device-resident throughput of the input-staging kernels (HBM-bound byte kernels) against the measured copy peak.
Usage: python tools/staging_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import refactored_orb_slam2_b200 as pkg
from test_staging import euroc_like_maps

peak = 6530.3
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


B, W, H = 320, 1241, 376                         # 446 MB of BGR input: larger than L2
src = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
dst = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
t = timed(lambda: pkg.staging.cvtColorGray_device(src.data_ptr(), 3, 0, W, H, dst.data_ptr(), batch=B, stream=stream))
byts = B * H * W * 4
print(json.dumps({"kernel": "k_cvt_gray<3>", "images": B, "shape": [H, W], "ms": round(t * 1e3, 4), "algorithmic_GBps": round(byts / t / 1e9, 1),
                  "frac_of_measured_hbm_peak": round(byts / t / 1e9 / peak, 3), "frames_per_s": round(B / t)}))
B, W, H = 512, 752, 480                          # EuRoC
mx, my = euroc_like_maps(W, H, 3)
dmx, dmy = torch.from_numpy(mx).to(dev), torch.from_numpy(my).to(dev)
src = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, device=dev)
dst = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
t = timed(lambda: pkg.staging.remap_device(src.data_ptr(), W, H, dmx.data_ptr(), dmy.data_ptr(), W, H, dst.data_ptr(), batch=B, stream=stream))
byts = B * H * W * 2                              # source read once + destination written once (the 2.9 MB of maps stay in L2)
print(json.dumps({"kernel": "k_remap", "images": B, "shape": [H, W], "ms": round(t * 1e3, 4), "algorithmic_GBps": round(byts / t / 1e9, 1),
                  "frac_of_measured_hbm_peak": round(byts / t / 1e9 / peak, 3), "frames_per_s": round(B / t)}))
