"""ORBmatcher Hamming searches: dense bank scan, candidate-list scan, DescriptorDistance and the
shard merge — indices and distances bit-exact against the oracle / numpy popcount."""
import numpy as np
import pytest

import gen_synth
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(built):
    import refactored_orb_slam2_b200 as p
    return p


def np_hamming_rows(q, bank):
    return np.bitwise_count(np.bitwise_xor(q[None, :], bank).view(np.uint64)).sum(axis=1)


@pytest.mark.parametrize("nb,nq", [(5000, 37), (100000, 2000), (777, 600), (130, 1), (20001, 1025), (128, 300), (129, 257)])
def test_dense_top2_vs_oracle(pkg, nb, nq):
    bank = gen_synth.descriptor_bank(nb, 4)
    if nb >= 4 * nq:
        q, bank, _ = gen_synth.queries_from_bank(bank, nq, 5)
    else:
        q = gen_synth.descriptor_bank(nq, 9)
    idx1, d1, d2 = pkg.DescriptorBank(bank, index_base=1000).top2(q)
    oi, o1, o2 = oracle.top2_dense(q, bank, 8)
    assert np.array_equal(idx1, oi + 1000) and np.array_equal(d1, o1) and np.array_equal(d2, o2)


def test_empty_bank_and_no_queries(pkg):
    q = gen_synth.descriptor_bank(5, 9)
    idx1, d1, d2 = pkg.DescriptorBank(np.zeros((0, 32), np.uint8)).top2(q)
    assert (idx1 == -1).all() and (d1 == 256).all() and (d2 == 256).all()       # ORBmatcher.cc:201-203 initial values
    idx1, d1, d2 = pkg.DescriptorBank(gen_synth.descriptor_bank(10, 1)).top2(np.zeros((0, 32), np.uint8))
    assert len(idx1) == 0


def test_ties_lowest_index_wins(pkg):
    """Duplicate bank rows: strict '<' keeps the first (ORBmatcher.cc:215-221), and the second-best
    equals the best when the best occurs twice."""
    base = gen_synth.descriptor_bank(50, 1)
    bank = np.tile(base, (40, 1))
    q = base[::3].copy()
    idx1, d1, d2 = pkg.DescriptorBank(bank).top2(q)
    assert np.array_equal(idx1, np.arange(0, 50, 3)) and (d1 == 0).all() and (d2 == 0).all()
    oi, o1, o2 = oracle.top2_dense(q, bank, 1)
    assert np.array_equal(idx1, oi) and np.array_equal(d2, o2)


def test_full_config4_exact(pkg):
    """BASELINE.json configs[3] at full size (2000 x 1M): ALL 2000 queries against the oracle's whole-bank scan (2 x 10^9
    pairs on every host thread), numpy popcount on a query subset as a third opinion, planted sources found, policy applied."""
    import os
    bank = gen_synth.descriptor_bank(1_000_000, 4)
    q, bank, src = gen_synth.queries_from_bank(bank, 2000, 5)
    m = pkg.ORBmatcher(0.7)
    db = pkg.DescriptorBank(bank)
    idx1, d1, d2 = db.top2(q)
    oi, o1, o2 = oracle.top2_dense(q, bank, os.cpu_count() or 1)
    assert np.array_equal(idx1, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2)
    for i in (0, 1, 7, 500, 1999):
        d = np_hamming_rows(q[i], bank)
        order = np.argsort(d, kind="stable")
        assert idx1[i] == order[0] and d1[i] == d[order[0]] and d2[i] == d[order[1]]
    assert (d1 <= np_hamming_rows_pairs(q, bank[src])).all()
    match, _, _ = m.SearchBank(q, db)
    ok = oracle.accept(d1, d2, 0.7, 50)
    assert np.array_equal(match, np.where(ok, idx1, -1))
    assert 100 < ok.sum() < 2000          # the decoys make the ratio test bite on both sides


def np_hamming_rows_pairs(a, b):
    return np.bitwise_count(np.bitwise_xor(a, b).view(np.uint64)).sum(axis=1)


@pytest.mark.parametrize("shards", [1, 2, 4, 8])
def test_sharded_bank_merge_is_exact(pkg, shards):
    """Row-sharded bank: per-shard top-2 records merged in shard order == one scan of the whole bank
    (the all-gather + merge of SURVEY.md section 8e, emulated on one GPU)."""
    torch = pytest.importorskip("torch")
    nb, nq = 40000, 700
    bank = gen_synth.descriptor_bank(nb, 4)
    q, bank, _ = gen_synth.queries_from_bank(bank, nq, 5)
    bank[nb // 2 + 5] = bank[3]                       # an exact duplicate across shards
    dq = torch.from_numpy(q).cuda()
    parts = torch.empty((shards, nq, 16), dtype=torch.uint8, device="cuda")
    bounds = np.linspace(0, nb, shards + 1).astype(int)
    banks = []
    for s in range(shards):
        banks.append(pkg.DescriptorBank(bank[bounds[s]:bounds[s + 1]], index_base=int(bounds[s])))
        banks[-1].top2_device(dq.data_ptr(), nq, parts[s].data_ptr())
    torch.cuda.synchronize()          # the per-bank scans ran on the banks' own streams
    out = torch.empty((nq, 16), dtype=torch.uint8, device="cuda")
    match = torch.empty(nq, dtype=torch.int64, device="cuda")
    pkg.merge_top2_device(parts.data_ptr(), shards, nq, out.data_ptr(), 0.7, 50, match.data_ptr())
    torch.cuda.synchronize()
    rec = out.cpu().numpy().view(pkg.TOP2_DTYPE).reshape(-1)
    oi, o1, o2 = oracle.top2_dense(q, bank, 8)
    assert np.array_equal(rec["idx"], oi) and np.array_equal(rec["d1"], o1) and np.array_equal(rec["d2"], o2)
    assert np.array_equal(match.cpu().numpy(), np.where(oracle.accept(o1, o2, 0.7, 50), oi, -1))


def test_candidate_list_search(pkg):
    rng = np.random.default_rng(3)
    train, q = gen_synth.descriptor_bank(3000, 2), gen_synth.descriptor_bank(500, 3)
    lens = rng.integers(0, 60, 500)
    lens[::50] = 0                                      # empty candidate lists
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.integers(0, 3000, off[-1]).astype(np.int32)
    cand[off[7]:off[8]] = cand[off[7]]                  # one list full of the same row: d2 == d1
    m = pkg.ORBmatcher(0.75)
    match, d1, d2 = m.SearchCandidates(q, train, off, cand)
    oi, o1, o2 = oracle.top2_csr(q, train, off, cand)
    assert np.array_equal(d1, o1) and np.array_equal(d2, o2)
    assert np.array_equal(match, np.where(oracle.accept(o1, o2, 0.75, 50), oi, -1))
    with pytest.raises(pkg.OrbGpuError):
        m.SearchCandidates(q[:1], train, np.array([0, 1], np.int32), np.array([3000], np.int32))


def test_descriptor_distance(pkg):
    a, b = gen_synth.descriptor_bank(1000, 5), gen_synth.descriptor_bank(1000, 6)
    d = pkg.ORBmatcher.DescriptorDistance(a, b)
    assert np.array_equal(d, np_hamming_rows_pairs(a, b))
    assert pkg.ORBmatcher.DescriptorDistance(a[0], a[0]) == 0
    assert pkg.ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256
    assert d[17] == oracle.hamming(a[17], b[17])


def test_self_matching_extracted_descriptors(pkg):
    """Extraction feeding matching on the device-resident path: a frame matched against itself."""
    img = gen_synth.texture(376, 1241, 2)
    k, d = pkg.ORBextractor(2000, 1.2, 8, 20, 7)(img)
    idx1, d1, d2 = pkg.DescriptorBank(d).top2(d)
    oi, o1, o2 = oracle.top2_dense(d, d, 4)
    assert np.array_equal(idx1, oi) and (d1 == 0).all() and np.array_equal(d2, o2)
