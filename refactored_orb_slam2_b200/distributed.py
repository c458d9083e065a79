"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Two paths shard:
  * frame / stereo-pair batches: contiguous ranges per rank, NO data-path collective
    (shard_range); results stay with the rank that produced them.
  * brute-force matching against a large keyframe bank: the bank is row-sharded in index order,
    queries are replicated, every rank scans its shard (liborbgpu.so), the per-shard {d1, d2, idx}
    records (16 B per query and rank) are exchanged and every rank folds them in rank order —
    identical to one sequential scan of the whole bank because ties go to the lowest index and each
    shard's top-2 contains any global top-2 member.  Two exchanges: "peer" (default on one node) —
    the library's own kernels store the records straight into every peer's HBM window over NVLink
    (CUDA IPC mappings, PeerExchange below) and the merge kernel waits on arrival flags, no
    collective call at all; "nccl" — one all_gather_into_tensor + the merge kernel.
torch.distributed is plumbing only (it carries the 64-byte IPC handles once, at set-up).
"""
import numpy as np

from ._lib import TOP2_DTYPE


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n units for `rank`; sizes differ by at most one, order preserved."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_records_host(parts):
    """Reference-order fold of [shards, nq] TOP2 records on the host (numpy).  Used where the merge has
    to run without a device (gloo tests); the product path uses merge_top2_device."""
    parts = np.asarray(parts).view(TOP2_DTYPE).reshape(len(parts), -1)
    d1 = np.full(parts.shape[1], 256, np.int32)
    d2 = np.full(parts.shape[1], 256, np.int32)
    idx = np.full(parts.shape[1], -1, np.int64)
    for p in parts:                      # ascending shard order == ascending bank index
        better = p["d1"] < d1
        d2 = np.where(better, np.minimum(d1, p["d2"]), np.minimum(d2, p["d1"]))
        idx = np.where(better, p["idx"], idx)
        d1 = np.where(better, p["d1"], d1)
    return idx, d1, d2


class PeerExchange:
    """This rank's record window + the mappings of its peers' windows (include/orbgpu.h, orbgpu_exchange_*)."""

    def __init__(self, rank, world, max_queries, device=-1, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from ._lib import check, lib
        self._h = C.c_void_p()
        handle = np.zeros(64, np.uint8)
        check(lib().orbgpu_exchange_create(rank, world, int(max_queries), device, C.byref(self._h), handle.ctypes.data_as(C.c_void_p)))
        self.rank, self.world, self.max_queries = rank, world, int(max_queries)
        if world > 1:
            mine = torch.from_numpy(handle).cuda()
            everyone = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(everyone, mine, group=group)
            handles = np.ascontiguousarray(everyone.cpu().numpy())
            check(lib().orbgpu_exchange_connect(self._h, handles.ctypes.data_as(C.c_void_p)))
            dist.barrier(group=group)          # every window is mapped everywhere before the first push

    def check(self):
        from ._lib import check, lib
        check(lib().orbgpu_exchange_check(self._h))

    def close(self):
        from ._lib import lib
        if getattr(self, "_h", None) is not None and self._h:
            lib().orbgpu_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedDescriptorBank:
    """Row shard of a global keyframe bank on this rank's GPU + the all-gather / merge step."""

    def __init__(self, global_rows, local_descriptors, rank, world, device=-1, group=None, exchange="peer", max_queries=8192):
        from .matcher import DescriptorBank
        self.rank, self.world, self.group = rank, world, group
        self.exchange = exchange
        self._peer = PeerExchange(rank, world, max_queries, device, group) if exchange == "peer" else None
        self.lo, self.hi = shard_range(global_rows, rank, world)
        local = np.ascontiguousarray(local_descriptors, np.uint8).reshape(-1, 32)
        if len(local) != self.hi - self.lo:
            raise ValueError("rank %d owns rows [%d, %d) but got %d rows" % (rank, self.lo, self.hi, len(local)))
        self.device = device
        self.bank = DescriptorBank(local, index_base=self.lo, device=device)

    def search(self, queries, nnratio=0.7, th_low=50):
        """Best / second-best of every query over the WHOLE bank and the accepted match index
        (ORBmatcher.cc:224-226), identical on every rank.  Returns (match, idx1, d1, d2) numpy arrays."""
        import torch
        import torch.distributed as dist

        from .matcher import merge_top2_device
        if getattr(self, "_stream", None) is None:
            self._stream = torch.cuda.Stream()      # explicit side stream (a NULL stream means "the bank's own")
        with torch.cuda.stream(self._stream):
            return self._search(queries, nnratio, th_low, torch, dist, merge_top2_device)

    def _search(self, queries, nnratio, th_low, torch, dist, merge_top2_device):
        q = torch.from_numpy(np.ascontiguousarray(queries, np.uint8).reshape(-1, 32)).cuda()
        nq = q.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        if self._peer is not None:
            import ctypes as C

            from ._lib import check, lib
            merged = torch.empty((nq, 16), dtype=torch.uint8, device="cuda")
            match = torch.empty(nq, dtype=torch.int64, device="cuda")
            check(lib().orbgpu_hamming_top2_sharded_device(self.bank._h, self._peer._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(merged.data_ptr()),
                                                           float(nnratio), int(th_low), C.c_void_p(match.data_ptr()), C.c_void_p(stream)))
            rec = merged.cpu().numpy().view(TOP2_DTYPE).reshape(-1)
            self._peer.check()
            return match.cpu().numpy(), rec["idx"].copy(), rec["d1"].copy(), rec["d2"].copy()
        local = torch.empty((nq, 16), dtype=torch.uint8, device="cuda")
        self.bank.top2_device(q.data_ptr(), nq, local.data_ptr(), stream=stream)
        if self.world > 1:
            parts = torch.empty((self.world * nq, 16), dtype=torch.uint8, device="cuda")   # rank-major
            dist.all_gather_into_tensor(parts, local, group=self.group)
        else:
            parts = local
        merged = torch.empty((nq, 16), dtype=torch.uint8, device="cuda")
        match = torch.empty(nq, dtype=torch.int64, device="cuda")
        merge_top2_device(parts.data_ptr(), self.world, nq, merged.data_ptr(), nnratio, th_low, match.data_ptr(),
                          device=self.device, stream=stream)
        rec = merged.cpu().numpy().view(TOP2_DTYPE).reshape(-1)
        return match.cpu().numpy(), rec["idx"].copy(), rec["d1"].copy(), rec["d2"].copy()
