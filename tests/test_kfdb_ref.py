"""SURVEY 8f row f4, second half, CPU side: KeyFrameDatabase::DetectLoopCandidates / DetectRelocalizationCandidates
(KeyFrameDatabase.cc:72-306).  The reference's OWN, unmodified KeyFrameDatabase.cc over its real DBoW2 scoring
(oracle/_ref/libref_kfdb.so) against the drop-in shim/KeyFrameDatabase.cc on the CPU oracle backend: candidate lists in order,
and the bookkeeping both leave on every KeyFrame (query ids, word counts, float scores bit for bit)."""
import numpy as np
import pytest

import kfdb_ref
import oracle
import voc_ref

pytestmark = pytest.mark.skipif(not (kfdb_ref.available("ref") and kfdb_ref.available("shim_cpu")),
                                reason="oracle/_ref/libref_kfdb.so absent and /root/reference not present")


@pytest.fixture(scope="module")
def vocab(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("kfdb") / "voc_k10_L4.txt")
    voc_ref.write_text_vocabulary(path, k=10, L=4, seed=3)
    return path


def drive(lib, bows, nwords, seed=0):
    """One scripted life of a database; returns everything observable."""
    rng = np.random.default_rng(seed)
    n = len(bows)
    log = []
    for k, (ids, w) in enumerate(bows):
        assert lib.new_keyframe(k, ids, w) == k             # mnId = index (so the id-0 quirk is exercised too)
    grng = np.random.default_rng(99)
    for k in range(n):
        lib.set_graph(k, *kfdb_ref.graph_of(n, k, grng))
    for k in range(n - 10):
        lib.add(k)

    def snap(tag, result):
        ints, sc = lib.state()
        log.append((tag, result.copy(), ints, sc.view(np.uint32).copy()))

    for q in (n - 1, n - 2, n - 5, 0, 57, n - 1):           # the last one repeats a query id: stale mnLoopQuery marks
        for min_score in (0.0, 0.02, 0.3):
            snap(("loop", q, min_score), lib.detect_loop(q, min_score))
    for f in range(4):
        ids, w = bows[int(rng.integers(0, n))]
        keep = rng.random(len(ids)) < 0.8
        snap(("reloc", f), lib.detect_reloc(1000 + f, ids[keep], w[keep] / w[keep].sum()))
    snap(("reloc_repeat", 0), lib.detect_reloc(1000, *bows[3]))      # the same frame id again: counts keep accumulating
    for k in (5, 6, 40, 41, 42):
        lib.erase(k)
    lib.add(40)                                             # re-added: goes to the END of its words' lists
    lib.add(7)                                              # added twice: counted twice per shared word
    for q in (n - 3, n - 4):
        snap(("loop_after_erase", q), lib.detect_loop(q, 0.01))
    snap(("reloc_after_erase",), lib.detect_reloc(2000, *bows[41]))
    lib.erase(7)                                            # removes the FIRST entry of the doubly added keyframe
    snap(("loop_after_erase2",), lib.detect_loop(n - 6, 0.0))
    lib.clear()
    snap(("loop_empty",), lib.detect_loop(n - 1, 0.0))
    for k in range(30):
        lib.add(k)
    snap(("loop_after_clear",), lib.detect_loop(n - 7, 0.0))
    snap(("reloc_no_words",), lib.detect_reloc(3000, np.zeros(0, np.uint32), np.zeros(0)))
    return log


def assert_same(a, b):
    assert len(a) == len(b)
    for (tag, ra, ia, sa), (_, rb, ib, sb) in zip(a, b):
        assert np.array_equal(ra, rb), (tag, ra, rb)
        assert np.array_equal(ia, ib), tag
        assert np.array_equal(sa, sb), tag


def test_shim_on_oracle_backend_equals_reference_keyframe_database(vocab):
    ref, shim = kfdb_ref.KfdbLib("ref", vocab), kfdb_ref.KfdbLib("shim_cpu", vocab)
    nwords = ref.L.voc_words(ref.voc)
    bows = kfdb_ref.bow_world(nwords, 140, 5)
    a, b = drive(ref, bows, nwords), drive(shim, bows, nwords)
    assert_same(a, b)
    found = [len(r) for tag, r, _, _ in a]
    assert max(found) >= 3 and sum(1 for f in found if f) >= 12            # the scenario does produce candidates
    ref.close(); shim.close()


def test_oracle_l1_score_equals_dbow2(vocab):
    """oracle.bow_query (what the device kernel is checked against) against TemplatedVocabulary::score -> L1Scoring::score of the
    reference's DBoW2, as doubles, bit for bit."""
    ref = kfdb_ref.KfdbLib("ref", vocab)
    bows = kfdb_ref.bow_world(ref.L.voc_words(ref.voc), 60, 8)
    common, first, score = oracle.bow_query(bows, bows[17])
    for k, kf in enumerate(bows):
        want = ref.voc_score(bows[17], kf)
        assert np.float64(want).view(np.uint64) == score[k].view(np.uint64), k
        shared = np.intersect1d(bows[17][0], kf[0])
        assert common[k] == len(shared) and first[k] == (shared[0] if len(shared) else 0xffffffff)
    assert score[17] > 0.99 and (score < 0.9).sum() > 40
    ref.close()
