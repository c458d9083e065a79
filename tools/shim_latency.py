"""Single stereo frame through the C++ drop-in a maintainer links (shim/shim_demo: ORBextractor::operator() x2 +
ComputeStereoMatchesGPU), with and without the host download of mvImagePyramid; prints shim_demo's JSON line."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from refactored_orb_slam2_b200 import synthetic

W, H = 1241, 376
left, right = synthetic.frame_batch(1, H, W, 100, stereo=True)
d = tempfile.mkdtemp()
left[0].tofile(os.path.join(d, "l.raw"))
right[0].tofile(os.path.join(d, "r.raw"))
out = subprocess.check_output([os.path.join(ROOT, "shim", "shim_demo"), os.path.join(d, "l.raw"), os.path.join(d, "r.raw"), str(W), str(H), "2000",
                               os.path.join(d, "out.bin"), "200"], text=True)
print([l for l in out.splitlines() if l.startswith("{")][0])
