"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Two paths shard:
  * frame / stereo-pair batches: contiguous ranges per rank, NO data-path collective
    (shard_range); results stay with the rank that produced them.
  * brute-force matching against a large keyframe bank: the bank is row-sharded in index order,
    queries are replicated, every rank scans its shard (liborbgpu.so), ONE all-gather exchanges the
    per-shard {d1, d2, idx} records (16 B per query and rank) and every rank folds them in rank order
    with the device merge kernel — identical to one sequential scan of the whole bank because ties
    go to the lowest index and each shard's top-2 contains any global top-2 member.
torch.distributed is plumbing only; the scan and the merge are the library's kernels.
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


class ShardedDescriptorBank:
    """Row shard of a global keyframe bank on this rank's GPU + the all-gather / merge step."""

    def __init__(self, global_rows, local_descriptors, rank, world, device=-1, group=None):
        from .matcher import DescriptorBank
        self.rank, self.world, self.group = rank, world, group
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
