"""Times the dense Hamming scan (2000 x 1M) for the ORBGPU_HAMMING_VARIANT of this process and checks 64 queries."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import gen_synth, oracle
import refactored_orb_slam2_b200 as pkg
bank = gen_synth.descriptor_bank(1_000_000, 4); q = gen_synth.descriptor_bank(2000, 5)
db = torch.from_numpy(bank).cuda(); dq = torch.from_numpy(q).cuda(); out = torch.empty((2000, 16), dtype=torch.uint8, device="cuda")
b = pkg.DescriptorBank(device_ptr=db.data_ptr(), rows=len(bank), keepalive=db)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for _ in range(3): b.top2_device(dq.data_ptr(), 2000, out.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10): b.top2_device(dq.data_ptr(), 2000, out.data_ptr(), stream=st.cuda_stream)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
rec = out.cpu().numpy().view(pkg.TOP2_DTYPE).reshape(-1)
i1, d1, d2 = oracle.top2_dense(q[:64], bank, 8)
ok = np.array_equal(rec["idx"][:64], i1) and np.array_equal(rec["d1"][:64], d1) and np.array_equal(rec["d2"][:64], d2)
print("variant %s: %.3f ms  %.0f Gpairs/s  parity(64 queries) %s" % (os.environ.get("ORBGPU_HAMMING_VARIANT", "default"), ms, 2e9 / ms / 1e6, ok))
