import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import gen_synth
import refactored_orb_slam2_b200 as pkg
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
img = gen_synth.texture(h, w, 1)
g = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_width=w, max_height=h)
k, d = g(img)
print("keypoints", len(k))
