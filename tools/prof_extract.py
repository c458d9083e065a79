"""Small, fixed workload for ncu: a few device-resident KITTI-shaped batches through the whole hot path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import gen_synth
import refactored_orb_slam2_b200 as pkg

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, H = 1241, 376
left, right = gen_synth.frame_batch(pairs, H, W, 100, stereo=True)
dev = torch.from_numpy(np.concatenate([left, right])).cuda()
ex = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=2 * pairs)
mb = float(np.float32(386.1448) / np.float32(718.856))
for _ in range(steps):
    ex.extract_batch_device(dev.data_ptr(), 2 * pairs, W, H)
    pkg._lib.check(pkg.lib().orbgpu_stereo_match_device(ex.handle, ex.handle, 0, pairs, pairs, 386.1448, mb))
ex.synchronize()
if len(sys.argv) > 3:   # also a dense Hamming scan
    bank = pkg.DescriptorBank(gen_synth.descriptor_bank(1_000_000, 4))
    q = gen_synth.descriptor_bank(2000, 5)
    for _ in range(2):
        bank.top2(q)
print("done")
