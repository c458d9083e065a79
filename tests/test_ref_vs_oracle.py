"""The reference's OWN ORBextractor.cc (compiled from /root/reference against the OpenCV API stand-in of
oracle/ref_shim, primitives = the cv2-pinned restatements, monotone allocator) against the oracle: keypoints
(position, order, size, angle, response, octave) and descriptors must be byte-identical.  This is what pins
the restated glue — cell loops, DivideNode / DistributeOctTree list discipline, IC_Angle, computeOrbDescriptor,
operator() — to the reference's code itself.  CPU only; skipped where the checker can be neither found nor built."""
import numpy as np
import pytest

import gen_synth
import oracle
import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_orb.so absent and /root/reference not present")

CASES = [(640, 480, 1000, 1, (1.2, 8, 20, 7)), (1241, 376, 2000, 2, (1.2, 8, 20, 7)), (752, 480, 1200, 1000, (1.2, 8, 20, 7)),
         (333, 251, 300, 5, (1.2, 8, 20, 7)), (1241, 376, 2000, 12, (1.2, 8, 12, 7)), (640, 480, 800, 21, (1.5, 4, 25, 10)),
         (400, 300, 4000, 8, (1.2, 8, 20, 7))]


@pytest.mark.parametrize("w,h,nf,seed,rest", CASES)
def test_reference_code_equals_oracle(w, h, nf, seed, rest):
    img = gen_synth.texture(h, w, seed)
    k, d = oracle.OracleExtractor(nf, *rest)(img)
    r = ref.ReferenceExtractor(nf, *rest)
    rk, rd = r(img)
    assert len(rk) == len(k)
    for f in k.dtype.names:
        assert np.array_equal(rk[f], k[f]), f
    assert np.array_equal(rd, d)


def test_reference_pyramid_and_repeat_calls():
    """Same handle, several frames (exercises the two-arena allocator), and mvImagePyramid bytes."""
    r = ref.ReferenceExtractor(1000)
    o = oracle.OracleExtractor(1000)
    for seed in (3, 4, 5, 6):
        img = gen_synth.texture(480, 640, seed)
        rk, rd = r(img)
        k, d = o(img)
        assert np.array_equal(rd, d) and np.array_equal(rk["x"], k["x"]) and np.array_equal(rk["angle"], k["angle"])
    for l in range(8):
        assert np.array_equal(r.level(l), o.level(l, 0))


def test_noise_image_many_ties():
    img = np.random.default_rng(5).integers(0, 256, (251, 333), dtype=np.uint8)
    k, d = oracle.OracleExtractor(300)(img)
    rk, rd = ref.ReferenceExtractor(300)(img)
    assert np.array_equal(rd, d) and np.array_equal(rk["x"], k["x"]) and np.array_equal(rk["y"], k["y"])


def _adversarial(kind, w, h, seed):
    if kind == "binary":
        return (np.random.default_rng(seed).integers(0, 2, (h, w)) * 255).astype(np.uint8)
    if kind == "dot_lattice":            # isolated dots every 2 px on a ramp: a fifth of level 0 survives NMS
        img = np.zeros((h, w), np.uint8)
        ii, jj = np.meshgrid(np.arange(0, w, 2) // 2, np.arange(0, h, 2) // 2)
        img[::2, ::2] = (50 + (ii * 24 + jj) % 192).astype(np.uint8)
        return img
    if kind == "low_contrast":           # grey levels 100..118 in 3x3 blocks: margins right at the two thresholds
        base = np.random.default_rng(seed).integers(100, 119, (h // 3 + 1, w // 3 + 1))
        return np.kron(base, np.ones((3, 3), np.int64))[:h, :w].astype(np.uint8)
    return gen_synth.texture(h, w, seed)


@pytest.mark.parametrize("kind,w,h,nf,seed", [("binary", 341, 263, 400, 9), ("dot_lattice", 341, 263, 500, 0),
                                              ("low_contrast", 310, 222, 400, 10), ("panorama", 1600, 260, 120, 12)])
def test_reference_code_equals_oracle_on_adversarial_images(kind, w, h, nf, seed):
    """The inputs of the GPU parity cases of the same names: the oracle they are checked against is itself the reference's
    code here (panorama: 7-10 octree roots per level, every level returns 4 x roots keypoints, more than its quota)."""
    img = _adversarial(kind, w, h, seed)
    o = oracle.OracleExtractor(nf)
    k, d = o(img)
    rk, rd = ref.ReferenceExtractor(nf)(img)
    assert len(rk) == len(k) > 0
    for f in k.dtype.names:
        assert np.array_equal(rk[f], k[f]), f
    assert np.array_equal(rd, d)
    if kind == "panorama":
        assert len(k) > 2 * nf
    if kind == "dot_lattice":
        assert len(o.stage(0, 0)) > (w - 38) * (h - 38) // 6


def test_stereo_batch_threads_agree_with_oracle_counts():
    left, right = gen_synth.frame_batch(6, 376, 1241, 50, stereo=True)
    params = (2000, 1.2, 8, 20, 7)
    cl, cr, m = ref.stereo_batch(params, left, right, 386.1448, np.float32(386.1448) / np.float32(718.856), 3)
    ol, orr, om = oracle.stereo_batch(params, left, right, 386.1448, np.float32(np.float32(386.1448) / np.float32(718.856)), 3)
    assert np.array_equal(cl, ol) and np.array_equal(cr, orr) and np.array_equal(m, om)


def test_stereo_batch_survives_many_calls():
    """bench.py's reference arm calls ref.stereo_batch once per step with fresh handles per worker thread; their arenas
    must go back to the free list (round 1: `ref arena: out of slabs` at call 17 of the driver's 25)."""
    rng = np.random.default_rng(0)
    left = rng.integers(0, 256, (128, 120, 160), dtype=np.uint8)
    right = rng.integers(0, 256, (128, 120, 160), dtype=np.uint8)
    first = None
    for _ in range(60):           # 60 calls x 32 threads x 2 handles x 2 slabs = 7680 slab requests against 1024 reserved
        out = ref.stereo_batch((300, 1.2, 8, 20, 7), left, right, 40.0, 0.1, 32)
        if first is None:
            first = out
        assert all(np.array_equal(a, b) for a, b in zip(first, out))
