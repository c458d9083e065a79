"""e2e through the host-buffer API for several chunk counts (ORBGPU_CHUNKS is read once per process)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gen_synth
import refactored_orb_slam2_b200 as pkg
W, H, pairs = 1241, 376, 160
nimg = 2 * pairs
left, right = gen_synth.frame_batch(pairs, H, W, 100, stereo=True)
host = torch.from_numpy(np.concatenate([left, right])).pin_memory()
ex = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=nimg)
lib = pkg.lib(); cap = ex.capacity; vp = ctypes.c_void_p
out_k = torch.empty((nimg, cap, 28), dtype=torch.uint8).pin_memory(); out_d = torch.empty((nimg, cap, 32), dtype=torch.uint8).pin_memory()
out_c = torch.empty(nimg, dtype=torch.int32).pin_memory(); out_u = torch.empty((pairs, cap), dtype=torch.float32).pin_memory(); out_z = torch.empty((pairs, cap), dtype=torch.float32).pin_memory()
mb = float(np.float32(386.1448) / np.float32(718.856))
def step():
    pkg._lib.check(lib.orbgpu_extract_batch(ex.handle, vp(host.data_ptr()), nimg, W, H, W, W * H, vp(out_k.data_ptr()), vp(out_d.data_ptr()), cap, vp(out_c.data_ptr())))
    pkg._lib.check(lib.orbgpu_stereo_match(ex.handle, ex.handle, 0, pairs, pairs, 386.1448, mb, vp(out_u.data_ptr()), vp(out_z.data_ptr()), cap))
for _ in range(3): step()
t = time.perf_counter()
for _ in range(10): step()
dt = (time.perf_counter() - t) / 10
print("chunks %s: %.3f ms/step %.0f frames/s" % (os.environ.get("ORBGPU_CHUNKS", "8"), dt * 1e3, nimg / dt))
