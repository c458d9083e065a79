"""SURVEY 8f row f4, second half, GPU side: the device-resident BoW database (orbgpu_bowdb_*, csrc/bowdb.cu) against the
oracle's restatement (pinned to DBoW2 in test_kfdb_ref.py), and the drop-in shim/KeyFrameDatabase.cc on liborbgpu.so against
the reference's OWN KeyFrameDatabase.cc."""
import numpy as np
import pytest

import kfdb_ref
import oracle
import voc_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(built):
    import refactored_orb_slam2_b200 as p
    return p


def test_device_database_equals_oracle_bit_for_bit(pkg):
    """3000 keyframes x ~270 words, queries with 0 .. 2000 words, erased slots, growth across reallocations: shared-word counts,
    first shared word and the L1 score (float64 bits) for every slot."""
    rng = np.random.default_rng(2)
    bows = kfdb_ref.bow_world(100000, 3000, 11, places=40)
    db = pkg.BowDatabase()
    for k, (ids, w) in enumerate(bows):
        assert db.add(ids, w) == k
    alive = np.ones(len(bows), np.uint8)
    for k in rng.choice(len(bows), 200, replace=False):
        db.erase(int(k))
        alive[k] = 0
    big = np.sort(rng.choice(100000, 2000, replace=False)).astype(np.uint32)
    bigw = rng.uniform(0.1, 3.0, 2000)
    queries = [bows[5], bows[1777], (big, bigw / bigw.sum()), (np.zeros(0, np.uint32), np.zeros(0)), (bows[9][0][:1], np.ones(1))]
    for q in queries:
        common, first, score = db.query(*q)
        oc, of, os_ = oracle.bow_query(bows, q, alive)
        assert np.array_equal(common, oc) and np.array_equal(first, of)
        assert np.array_equal(score.view(np.uint64), os_.view(np.uint64))
    assert db.query(*bows[5])[0].max() == len(bows[5][0])
    db.clear()
    assert len(db) == 0 and db.add(*bows[0]) == 0
    with pytest.raises(pkg.OrbGpuError):
        db.add(np.array([5, 3], np.uint32), np.ones(2))                    # not ascending


@pytest.mark.skipif(not (kfdb_ref.available("ref") and kfdb_ref.available("shim_gpu")), reason="oracle/_ref/lib*_kfdb.so absent")
def test_shim_on_gpu_equals_reference_keyframe_database(pkg, tmp_path):
    from test_kfdb_ref import assert_same, drive
    path = str(tmp_path / "voc_k10_L4.txt")
    voc_ref.write_text_vocabulary(path, k=10, L=4, seed=3)
    ref, shim = kfdb_ref.KfdbLib("ref", path), kfdb_ref.KfdbLib("shim_gpu", path)
    nwords = ref.L.voc_words(ref.voc)
    for seed in (5, 6):
        bows = kfdb_ref.bow_world(nwords, 140, seed)
        assert_same(drive(ref, bows, nwords), drive(shim, bows, nwords))
        ref.close(); shim.close()
        ref, shim = kfdb_ref.KfdbLib("ref", path), kfdb_ref.KfdbLib("shim_gpu", path)
    ref.close(); shim.close()
