"""Generates tests/golden/orb_320x240.npz with the cv2-driven restatement (oracle/oracle_cv2.py):
real OpenCV resize / FAST / GaussianBlur / fastAtan2 + libm cosf/sinf on this container
(cv2 4.13.0, glibc 2.39).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import gen_synth
from oracle_cv2 import Cv2Extractor

if __name__ == "__main__":
    img = gen_synth.texture(240, 320, 42)
    params = (400, 1.2, 8, 20, 7)
    ex = Cv2Extractor(*params)
    k, d = ex(img)
    levels = {"level%d" % l: ex.level(l).copy() for l in range(8)}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_320x240.npz")
    np.savez_compressed(out, image=img, params=np.array(params, np.float64), keypoints=k, descriptors=d, **levels)
    print("wrote", out, len(k), "keypoints", os.path.getsize(out), "bytes")
