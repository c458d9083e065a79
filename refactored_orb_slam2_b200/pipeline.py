"""StereoBatchPipeline — throughput front-end for streams of stereo batches in HOST memory.

A single orbgpu_extract_batch call already overlaps the H2D copy, the kernels and the D2H copy of its own
chunks; what it cannot hide is the first chunk's upload and the last chunk's download.  This class keeps
`depth` extractor handles (each = its own CUDA streams and device buffers, see include/orbgpu.h "Threading")
in flight from ONE host thread through the split entry points (orbgpu_extract_batch_async +
orbgpu_stereo_match_async, then orbgpu_wait): batch k+1 is fully enqueued — its upload starts at once on its own
copy stream — before batch k is awaited, so uploads, kernels and downloads of neighbouring batches overlap in a
fixed, reproducible order.  Results are byte-identical to sequential orbgpu_extract_batch / orbgpu_stereo_match calls.
"""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib
from .extractor import ORBextractor


class _Slot:
    def __init__(self, params, width, height, pairs, device, pin):
        self.ex = ORBextractor(*params, max_width=width, max_height=height, max_batch=2 * pairs, device=device)
        cap = self.ex.capacity
        self.cap = cap
        self.kps = np.empty((2 * pairs, cap), KP_DTYPE)
        self.desc = np.empty((2 * pairs, cap, 32), np.uint8)
        self.counts = np.zeros(2 * pairs, np.int32)
        self.u_right = np.empty((pairs, cap), np.float32)
        self.depth = np.empty((pairs, cap), np.float32)
        self.busy = False
        if pin:
            try:
                import torch
                for name in ("kps", "desc", "counts", "u_right", "depth"):
                    a = getattr(self, name)
                    t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
                    setattr(self, name, t.numpy().view(a.dtype).reshape(a.shape))
                    setattr(self, "_keep_" + name, t)
            except Exception:
                pass


class StereoBatchPipeline:
    """submit(images) with images = uint8 [2*pairs, H, W] (left images, then right images, pinned for real overlap);
    results come back in submission order from result() as (keypoints, descriptors, counts, mvuRight, mvDepth) views
    into the slot's buffers — consume them before `depth` further submits.  At most `depth` batches may be outstanding."""

    def __init__(self, params, width, height, pairs, mbf, mb, depth=2, device=-1, pin_outputs=True):
        self.width, self.height, self.pairs = int(width), int(height), int(pairs)
        self.mbf, self.mb = float(mbf), float(mb)
        self.depth = int(depth)
        self._slots = [_Slot(params, width, height, pairs, device, pin_outputs) for _ in range(self.depth)]
        self._submitted = 0
        self._returned = 0

    def submit(self, images):
        """Enqueues one batch; `images` must stay alive and unchanged until its result has been returned."""
        if self._submitted - self._returned >= self.depth:
            raise RuntimeError("%d batches already outstanding: call result() first" % self.depth)
        s = self._slots[self._submitted % self.depth]
        addr = images.data_ptr() if hasattr(images, "data_ptr") else images.ctypes.data
        n, vp, L = 2 * self.pairs, C.c_void_p, lib()
        check(L.orbgpu_extract_batch_async(s.ex.handle, vp(addr), n, self.width, self.height, self.width, self.width * self.height,
                                           vp(s.kps.ctypes.data), vp(s.desc.ctypes.data), s.cap, vp(s.counts.ctypes.data)))
        check(L.orbgpu_stereo_match_async(s.ex.handle, s.ex.handle, 0, self.pairs, self.pairs, self.mbf, self.mb,
                                          vp(s.u_right.ctypes.data), vp(s.depth.ctypes.data), s.cap))
        s.busy = True
        self._submitted += 1

    def result(self):
        """Blocks for the oldest outstanding batch."""
        if self._returned >= self._submitted:
            raise RuntimeError("no batch outstanding")
        s = self._slots[self._returned % self.depth]
        self._returned += 1
        s.busy = False
        check(lib().orbgpu_wait(s.ex.handle))
        return s.kps, s.desc, s.counts, s.u_right, s.depth

    def close(self):
        for s in self._slots:
            if s.busy:
                lib().orbgpu_wait(s.ex.handle)
                s.busy = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
