"""StereoBatchPipeline — throughput front-end for streams of stereo batches in HOST memory.

A single orbgpu_extract_batch call already overlaps the H2D copy, the kernels and the D2H copy of its own
chunks; what it cannot hide is the first chunk's upload and the last chunk's download.  This class keeps
`depth` extractor handles (each = its own CUDA streams and device buffers, see include/orbgpu.h "Threading")
busy from `depth` host threads, so that batch k+1 uploads while batch k computes and batch k-1 downloads.
Every batch still goes through the public host-buffer entry points (orbgpu_extract_batch +
orbgpu_stereo_match); results are byte-identical to sequential calls.
"""
import ctypes as C
import queue
import threading

import numpy as np

from ._lib import KP_DTYPE, check, lib
from .extractor import ORBextractor


class _Slot:
    def __init__(self, params, width, height, pairs, device):
        self.ex = ORBextractor(*params, max_width=width, max_height=height, max_batch=2 * pairs, device=device)
        cap = self.ex.capacity
        self.cap = cap
        self.kps = np.empty((2 * pairs, cap), KP_DTYPE)
        self.desc = np.empty((2 * pairs, cap, 32), np.uint8)
        self.counts = np.zeros(2 * pairs, np.int32)
        self.u_right = np.empty((pairs, cap), np.float32)
        self.depth = np.empty((pairs, cap), np.float32)


class StereoBatchPipeline:
    """submit(images) with images = uint8 [2*pairs, H, W] (left images, then right images, ideally pinned);
    results come back in submission order from results() / drain() as (keypoints, descriptors, counts,
    mvuRight, mvDepth) views into the slot's buffers — consume them before 2*depth further submits."""

    def __init__(self, params, width, height, pairs, mbf, mb, depth=2, device=-1, pin_outputs=True):
        self.width, self.height, self.pairs = int(width), int(height), int(pairs)
        self.mbf, self.mb = float(mbf), float(mb)
        self.depth = int(depth)
        # two result buffers per worker so that a returned result stays valid while the worker runs its next batch
        self._slots = [[_Slot(params, width, height, pairs, device) for _ in range(2)] for _ in range(self.depth)]
        if pin_outputs:
            try:
                import torch
                for pair in self._slots:
                    for s in pair:
                        for name in ("kps", "desc", "counts", "u_right", "depth"):
                            a = getattr(s, name)
                            t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
                            setattr(s, name, t.numpy().view(a.dtype).reshape(a.shape))
                            setattr(s, "_keep_" + name, t)
            except Exception:
                pass
        self._in = [queue.Queue() for _ in range(self.depth)]
        self._out = [queue.Queue() for _ in range(self.depth)]
        self._submitted = 0
        self._returned = 0
        self._threads = [threading.Thread(target=self._work, args=(i,), daemon=True) for i in range(self.depth)]
        for t in self._threads:
            t.start()

    def _work(self, i):
        turn = 0
        while True:
            job = self._in[i].get()
            if job is None:
                return
            s = self._slots[i][turn]
            turn ^= 1
            try:
                n = 2 * self.pairs
                vp = C.c_void_p
                check(lib().orbgpu_extract_batch(s.ex.handle, vp(job), n, self.width, self.height, self.width,
                                                 self.width * self.height, vp(s.kps.ctypes.data), vp(s.desc.ctypes.data),
                                                 s.cap, vp(s.counts.ctypes.data)))
                check(lib().orbgpu_stereo_match(s.ex.handle, s.ex.handle, 0, self.pairs, self.pairs, self.mbf, self.mb,
                                                vp(s.u_right.ctypes.data), vp(s.depth.ctypes.data), s.cap))
                self._out[i].put((s.kps, s.desc, s.counts, s.u_right, s.depth))
            except Exception as e:      # surfaced to the caller of results()
                self._out[i].put(e)

    def submit(self, images):
        """Queues one batch; `images` must stay alive and unchanged until its result has been returned."""
        addr = images.data_ptr() if hasattr(images, "data_ptr") else images.ctypes.data
        self._in[self._submitted % self.depth].put(addr)
        self._submitted += 1

    def result(self):
        """Blocks for the oldest outstanding batch."""
        if self._returned >= self._submitted:
            raise RuntimeError("no batch outstanding")
        r = self._out[self._returned % self.depth].get()
        self._returned += 1
        if isinstance(r, Exception):
            raise r
        return r

    def close(self):
        for q in self._in:
            q.put(None)
        for t in self._threads:
            t.join(timeout=10)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
