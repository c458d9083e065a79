"""The C++ oracle against the committed golden vectors (tests/golden/orb_320x240.npz, produced by
the cv2-driven restatement: real OpenCV primitives, see tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np

import oracle


def test_oracle_reproduces_golden_vectors():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "orb_320x240.npz"))
    nf, sf, nl, ini, mn = z["params"]
    e = oracle.OracleExtractor(int(nf), float(sf), int(nl), int(ini), int(mn))
    k, d = e(z["image"])
    assert len(k) == len(z["keypoints"]) == 395
    for f in k.dtype.names:
        assert np.array_equal(k[f], z["keypoints"][f]), f
    assert np.array_equal(d, z["descriptors"])
    for l in range(int(nl)):
        assert np.array_equal(e.level(l, 0), z["level%d" % l])
