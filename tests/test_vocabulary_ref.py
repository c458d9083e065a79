"""SURVEY 8f row f2, CPU side: the oracle's restatement of the DBoW2 tree descent and shim/ORBVocabularyGPU.h's bookkeeping
against the reference's OWN ORBVocabulary.cc + vendored DBoW2 (oracle/_ref/libref_voc.so), on synthetic vocabularies written
in the reference's text format and read back by the reference's own loader."""
import numpy as np
import pytest

import oracle
import voc_ref

pytestmark = pytest.mark.skipif(not (voc_ref.available("ref") and voc_ref.available("shim_cpu")),
                                reason="oracle/_ref/libref_voc.so absent and /root/reference not present")


@pytest.fixture(scope="module")
def vocab(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("voc") / "voc_k10_L4.txt")
    n = voc_ref.write_text_vocabulary(path, k=10, L=4, seed=1)
    return path, n


def test_reference_loader_reads_the_synthetic_vocabulary(vocab):
    path, n = vocab
    v = voc_ref.VocLib("ref", path)
    assert v.nodes == n and v.depth == 4 and 5000 < v.words < v.nodes
    start, kids, desc, word, weight = v.export()
    assert start[-1] == n - 1 and (np.diff(start) <= 10).all()
    assert (weight[word >= 0] == 0).sum() > 10          # stopped words exist
    v.close()


@pytest.mark.parametrize("levelsup", [4, 2, 1, 0, 6])
def test_oracle_descent_equals_reference_descent(vocab, levelsup):
    path, _ = vocab
    v = voc_ref.VocLib("ref", path)
    start, kids, desc, word, weight = v.export()
    feats = voc_ref.features_near_words(desc, word, np.random.default_rng(levelsup), 3000)
    rw, rweight, rnid = v.transform_each(feats, levelsup)
    ow, on = oracle.voc_transform(start, kids, desc, word, v.depth, feats, levelsup)
    assert np.array_equal(rw, ow)
    touched = rnid != 0xffffffff                         # the reference leaves *nid untouched for leaves above the level
    assert np.array_equal(rnid[touched].astype(np.int32), on[touched])
    if levelsup in (0, 1):            # leaves at levels 2 and 3 exist; nid_level = 4 or 3
        assert (~touched).sum() > 0
        leaf_of_word = np.full(v.words, -1, np.int64)
        leaf_of_word[word[word >= 0]] = np.nonzero(word >= 0)[0]
        assert np.array_equal(on[~touched], leaf_of_word[ow[~touched]])       # the documented rule: the leaf's own node id
    assert len(np.unique(rw)) > 1000
    v.close()


@pytest.mark.parametrize("levelsup", [4, 2])
def test_shim_vocabulary_on_oracle_backend_equals_reference_transform(vocab, levelsup):
    """BowVector (ids, and tf-idf values as doubles, bit for bit) and FeatureVector through the virtual transform() that
    Frame::ComputeBoW calls: reference vs shim/ORBVocabularyGPU.h on the CPU oracle backend.  Features whose FeatureVector node
    the reference leaves uninitialised are kept out (the reference's own answer is undefined for them)."""
    path, _ = vocab
    r, s = voc_ref.VocLib("ref", path), voc_ref.VocLib("shim_cpu", path)
    start, kids, desc, word, weight = r.export()
    feats = voc_ref.features_near_words(desc, word, np.random.default_rng(9), 2500)
    _, _, rnid = r.transform_each(feats, levelsup)
    feats = feats[rnid != 0xffffffff]
    assert len(feats) > 2000
    a, b = r.transform(feats, levelsup), s.transform(feats, levelsup)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    assert len(a["bow_ids"]) > 800 and abs(a["bow_vals"].sum() - 1.0) < 1e-9        # L1-normalised
    r.close(); s.close()
