"""SURVEY 8f row f2 on the GPU: the DBoW2 tree descent (csrc/search.cu, k_voc_transform) through the C ABI against the
oracle, and shim/ORBVocabularyGPU.h on liborbgpu.so against the reference's own ORBVocabulary + DBoW2.  Bit-exact."""
import numpy as np
import pytest

import oracle
import voc_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import refactored_orb_slam2_b200 as p
    return p


def synthetic_tree(k, L, seed, shallow=0.05):
    """Flat arrays of a k-ary tree (same construction as voc_ref.write_text_vocabulary, without the text round trip)."""
    rng = np.random.default_rng(seed)
    desc, start, kids, word = [np.zeros(32, np.uint8)], [], [], [-1]
    frontier, level_of, nwords = [(0, rng.integers(0, 2, 256, dtype=np.uint8), 0)], {0: 0}, 0
    children_of = {}
    while frontier:
        new = []
        for pid, bits, level in frontier:
            children_of[pid] = []
            for _ in range(k):
                child = bits.copy()
                child[rng.choice(256, 24, replace=False)] ^= 1
                nid = len(desc)
                desc.append(np.packbits(child))
                leaf = level + 1 == L or (level + 1 >= 2 and rng.random() < shallow)
                word.append(nwords if leaf else -1)
                nwords += int(leaf)
                children_of[pid].append(nid)
                if not leaf:
                    new.append((nid, child, level + 1))
        frontier = new
    n = len(desc)
    for i in range(n):
        start.append(len(kids))
        kids.extend(children_of.get(i, []))
    start.append(len(kids))
    return np.array(start, np.int32), np.array(kids, np.int32), np.stack(desc), np.array(word, np.int32)


@pytest.mark.parametrize("k,L,levelsup", [(10, 4, 4), (10, 4, 2), (10, 4, 1), (10, 4, 0), (10, 5, 4), (7, 3, 1), (20, 3, 2), (3, 6, 4)])
def test_descent_vs_oracle(pkg, k, L, levelsup):
    start, kids, desc, word = synthetic_tree(k, L, seed=k * 10 + L)
    voc = pkg.Vocabulary(start, kids, desc, word, L)
    feats = voc_ref.features_near_words(desc, word, np.random.default_rng(3), 5000)
    feats[:40] = desc[kids[:40]]                         # exact hits and, below, exact ties between siblings
    gw, gn = voc.transform(feats, levelsup)
    ow, on = oracle.voc_transform(start, kids, desc, word, L, feats, levelsup)
    assert np.array_equal(gw, ow) and np.array_equal(gn, on)
    assert len(np.unique(gw)) > min(500, (word >= 0).sum() // 4)


def test_ties_between_children_go_to_the_first(pkg):
    start, kids, desc, word = synthetic_tree(10, 3, seed=5, shallow=0.0)
    desc = desc.copy()
    desc[kids[start[0]:start[1]]] = desc[kids[0]]         # all ten children of the root identical: strict '<' keeps the first
    voc = pkg.Vocabulary(start, kids, desc, word, 3)
    feats = np.random.default_rng(0).integers(0, 256, (200, 32), dtype=np.uint8)
    gw, gn = voc.transform(feats, 2)
    assert (gn == kids[0]).all()
    ow, on = oracle.voc_transform(start, kids, desc, word, 3, feats, 2)
    assert np.array_equal(gw, ow) and np.array_equal(gn, on)


def test_device_resident_features_and_edge_cases(pkg):
    from refactored_orb_slam2_b200._lib import KP_DTYPE
    start, kids, desc, word = synthetic_tree(10, 4, seed=2)
    voc = pkg.Vocabulary(start, kids, desc, word, 4)
    feats = voc_ref.features_near_words(desc, word, np.random.default_rng(1), 2000)
    kp = np.zeros(len(feats), KP_DTYPE)
    kp["x"], kp["y"] = np.random.default_rng(2).uniform(20, 600, (2, len(feats)))
    fs = pkg.FrameFeatures(kp, feats, None, pkg.grid_geometry(0.0, 640.0, 0.0, 480.0))
    a, b = voc.transform(fs, 4), voc.transform(feats, 4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    w, n = voc.transform(np.zeros((0, 32), np.uint8), 4)
    assert len(w) == 0
    with pytest.raises(pkg.OrbGpuError):                  # a child that precedes its parent: the descent could loop
        bad = kids.copy(); bad[5] = 0
        pkg.Vocabulary(start, bad, desc, word, 4)


@pytest.mark.skipif(not (voc_ref.available("ref") and voc_ref.available("shim_gpu")), reason="oracle/_ref checker libraries not shipped with this snapshot")
@pytest.mark.parametrize("levelsup", [4, 2])
def test_shim_vocabulary_on_gpu_equals_reference_transform(pkg, tmp_path, levelsup):
    """Frame::ComputeBoW's call — the virtual transform(features, BowVector, FeatureVector, levelsup) — on the reference's own
    ORBVocabulary + DBoW2 and on shim/ORBVocabularyGPU.h over liborbgpu.so: BowVector ids and tf-idf doubles, FeatureVector."""
    path = str(tmp_path / "voc.txt")
    voc_ref.write_text_vocabulary(path, k=10, L=4, seed=1)
    r, s = voc_ref.VocLib("ref", path), voc_ref.VocLib("shim_gpu", path)
    start, kids, desc, word, weight = r.export()
    feats = voc_ref.features_near_words(desc, word, np.random.default_rng(9), 2500)
    rw, _, rnid = r.transform_each(feats, levelsup)
    gw, gn = pkg.Vocabulary(start, kids, desc, word, r.depth).transform(feats, levelsup)
    assert np.array_equal(gw, rw)
    touched = rnid != 0xffffffff
    assert np.array_equal(gn[touched], rnid[touched].astype(np.int32))
    feats = feats[touched]
    a, b = r.transform(feats, levelsup), s.transform(feats, levelsup)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    r.close(); s.close()
