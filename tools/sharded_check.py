"""Multi-GPU check of the row-sharded bank (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/sharded_check.py
Every rank scans its shard; the records are exchanged (a) by the library's peer-memory kernels and (b) by one NCCL
all-gather; both must equal a single-GPU scan of the WHOLE bank on every rank, step after step (epochs alternate windows).
Prints one JSON line from rank 0 with the device time per step of both exchanges."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import refactored_orb_slam2_b200 as pkg
from refactored_orb_slam2_b200 import distributed as D
from refactored_orb_slam2_b200 import synthetic

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rows = int(os.environ.get("CHECK_ROWS", 200003))
bank = synthetic.descriptor_bank(rows, 4)
lo, hi = D.shard_range(rows, rank, world)
whole = pkg.DescriptorBank(bank, device=local_rank)
peer = D.ShardedDescriptorBank(rows, bank[lo:hi], rank, world, device=local_rank, exchange="peer", max_queries=4096)
nccl = D.ShardedDescriptorBank(rows, bank[lo:hi], rank, world, device=local_rank, exchange="nccl")
bad = 0
for step in range(24):
    nq = [2000, 37, 1, 4096, 513][step % 5]
    q, _, _ = synthetic.queries_from_bank(bank, nq, 100 + step)
    idx, d1, d2 = whole.top2(q)
    want = np.where((d1 <= 50) & (d1.astype(np.float32) < np.float32(0.7) * d2.astype(np.float32)), idx, -1)
    for name, sb in (("peer", peer), ("nccl", nccl)):
        m, i, a, b = sb.search(q, 0.7, 50)
        if not (np.array_equal(i, idx) and np.array_equal(a, d1) and np.array_equal(b, d2) and np.array_equal(m, want)):
            bad += 1
            print("rank %d step %d %s MISMATCH" % (rank, step, name), file=sys.stderr)
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
# timing: device-resident loop of both exchanges on BASELINE config 4's shape (2000 queries), shard of a 1M-row bank
import ctypes as C
from refactored_orb_slam2_b200._lib import check, lib
big = np.random.default_rng(4 + rank).integers(0, 256, size=(1_000_000 // world, 32), dtype=np.uint8)
dbig = torch.from_numpy(big).cuda()
bk = pkg.DescriptorBank(device_ptr=dbig.data_ptr(), rows=len(big), index_base=rank * len(big), device=local_rank, keepalive=dbig)
ex = D.PeerExchange(rank, world, 2000, local_rank)
dq = torch.from_numpy(np.random.default_rng(5).integers(0, 256, size=(2000, 32), dtype=np.uint8)).cuda()
local = torch.empty((2000, 16), dtype=torch.uint8, device="cuda")
gathered = torch.empty((world * 2000, 16), dtype=torch.uint8, device="cuda")
merged = torch.empty((2000, 16), dtype=torch.uint8, device="cuda")
match = torch.empty(2000, dtype=torch.int64, device="cuda")
cur = torch.cuda.Stream()
torch.cuda.set_stream(cur)


def step_peer():
    check(lib().orbgpu_hamming_top2_sharded_device(bk._h, ex._h, C.c_void_p(dq.data_ptr()), 2000, C.c_void_p(merged.data_ptr()), 0.7, 50,
                                                   C.c_void_p(match.data_ptr()), C.c_void_p(cur.cuda_stream)))


def step_nccl():
    bk.top2_device(dq.data_ptr(), 2000, local.data_ptr(), stream=cur.cuda_stream)
    dist.all_gather_into_tensor(gathered, local)
    pkg.merge_top2_device(gathered.data_ptr(), world, 2000, merged.data_ptr(), 0.7, 50, match.data_ptr(), device=local_rank, stream=cur.cuda_stream)


res = {}
keep = {}
for name, fn in (("peer", step_peer), ("nccl", step_nccl)):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(cur)
    for _ in range(30):
        fn()
    b.record(cur)
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / 30], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[name + "_ms_per_step"] = round(float(ms.item()), 4)
    keep[name] = merged.cpu().numpy().copy()
ex.check()
res["peer_equals_nccl"] = bool(np.array_equal(keep["peer"], keep["nccl"]))
if rank == 0:
    res.update(world=world, mismatches=int(t.item()), rows_checked=rows,
               gpairs_per_s_peer=round(2000 * 1_000_000 / (res["peer_ms_per_step"] * 1e-3) / 1e9, 1),
               gpairs_per_s_nccl=round(2000 * 1_000_000 / (res["nccl_ms_per_step"] * 1e-3) / 1e9, 1))
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
