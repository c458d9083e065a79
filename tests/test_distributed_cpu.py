"""world_size-2 (and 3) gloo runs on CPU: the host-side logic of the multi-GPU paths — contiguous
sharding, index bases, the all-gather layout and the rank-order fold of per-shard top-2 records —
checked against the oracle's single scan.  The per-shard scans are played by the oracle here (test
infrastructure); on the GPU box the same fold runs as a kernel (tests/test_gpu_hamming.py)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gen_synth
    import oracle
    from refactored_orb_slam2_b200._lib import TOP2_DTYPE
    from refactored_orb_slam2_b200.distributed import merge_records_host, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb, nq = 30001, 257
    bank = gen_synth.descriptor_bank(nb, 4)
    q, bank, _ = gen_synth.queries_from_bank(bank, nq, 5)
    bank[20000] = bank[17]                                   # duplicate rows in different shards: lowest index must win
    lo, hi = shard_range(nb, rank, world)
    i1, d1, d2 = oracle.top2_dense(q, bank[lo:hi], 1)        # this rank's shard scan
    local = np.zeros(nq, TOP2_DTYPE)
    local["d1"], local["d2"], local["idx"] = d1, d2, np.where(i1 >= 0, i1 + lo, -1)
    lt = torch.from_numpy(local.view(np.uint8).reshape(nq, 16).copy())
    parts = torch.empty((world * nq, 16), dtype=torch.uint8)      # rank-major concatenation
    dist.all_gather_into_tensor(parts, lt)
    idx, m1, m2 = merge_records_host(parts.numpy().reshape(world, nq, 16))
    gi, g1, g2 = oracle.top2_dense(q, bank, 1)
    ok = np.array_equal(idx, gi) and np.array_equal(m1, g1) and np.array_equal(m2, g2)
    # frame sharding: every frame owned exactly once, in order
    owned = torch.zeros(101, dtype=torch.int32)
    a, b = shard_range(101, rank, world)
    owned[a:b] = 1
    dist.all_reduce(owned)
    ok = ok and bool((owned == 1).all())
    open(os.path.join(out_dir, "rank%d" % rank), "w").write("ok" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bank_merge_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def test_shard_range_properties():
    from refactored_orb_slam2_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 4096, 1_000_000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
