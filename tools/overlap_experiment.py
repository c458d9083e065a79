"""Experiment: one handle on the whole batch vs several handles on sub-batches launched back to back (own streams)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gen_synth
import refactored_orb_slam2_b200 as pkg
W, H = 1241, 376
pairs = 160
left, right = gen_synth.frame_batch(pairs, H, W, 100, stereo=True)
dev = torch.from_numpy(np.concatenate([left, right])).cuda()
mb = float(np.float32(386.1448) / np.float32(718.856))
lib = pkg.lib()
def run(nh, steps=10):
    n = 2 * pairs // nh
    hs = [pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=n) for _ in range(nh)]
    pp = pairs // nh
    bufs = [torch.cat([dev[k * pp:(k + 1) * pp], dev[pairs + k * pp: pairs + (k + 1) * pp]]).contiguous() for k in range(nh)]
    def step():
        for k, h in enumerate(hs):
            h.extract_batch_device(bufs[k].data_ptr(), n, W, H)
        for k, h in enumerate(hs):
            pkg._lib.check(lib.orbgpu_stereo_match_device(h.handle, h.handle, 0, pp, pp, 386.1448, mb))
    for _ in range(3): step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(steps): step()
    for h in hs: h.synchronize()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / steps
    print("handles %d FAST_CTAS %s: %.3f ms/step  %.0f frames/s" % (nh, os.environ.get("ORBGPU_FAST_CTAS", "4"), dt * 1e3, 2 * pairs / dt), flush=True)
run(int(sys.argv[1]))
