"""Parity tests proper: the CUDA extractor (through the C ABI) against the CPU oracle on the same
seeded inputs — every stage and the final cv::KeyPoint records / descriptors, bit-exact.
Float fields (angle, pt, size, response) are compared bit-exact as well; north_star's tolerance
(<= 1e-5 relative on angle) is slack we do not use."""
import numpy as np
import pytest

import gen_synth
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(built):
    import refactored_orb_slam2_b200 as p
    return p


def assert_same_keypoints(gk, gd, ok, od):
    assert len(gk) == len(ok)
    for f in ok.dtype.names:
        assert np.array_equal(gk[f], ok[f]), "keypoint field %s" % f
    assert np.array_equal(gd, od), "descriptors"


CONFIGS = [
    ("C1_tum1", 640, 480, 1000, 1, "texture"),          # BASELINE.json configs[0]
    ("C2_kitti", 1241, 376, 2000, 2, "texture"),        # configs[1]
    ("C3_euroc", 752, 480, 1200, 1000, "texture"),      # configs[2]
    ("noise", 333, 251, 300, 5, "noise"),               # dense corners: iniTh cells, big candidate lists
    ("odd_small", 301, 227, 150, 6, "texture"),         # top level 84x63: one cell row, 1 root, tiny quotas
    ("flat", 320, 240, 500, 0, "flat"),                 # no corners anywhere: empty output
    ("few", 400, 300, 4000, 8, "texture"),              # quota far above the candidate count
    ("binary", 341, 263, 400, 9, "binary"),             # 0 / 255 only: arc margins at the ends of the packed 16-bit range
    ("low_contrast", 310, 222, 400, 10, "low_contrast"),  # grey levels 100..118: margins sit right at minThFAST / iniThFAST
    ("panorama", 1600, 260, 120, 12, "texture"),        # 7-10 octree roots per level: every level returns 4 x roots > its quota
    ("dot_lattice", 341, 263, 500, 0, "dot_lattice"),   # isolated dots every 2 px on a ramp: 21 % of level 0 survives NMS
]


def make_image(w, h, seed, kind):
    if kind == "texture":
        return gen_synth.texture(h, w, seed)
    if kind == "noise":
        return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "binary":
        return (np.random.default_rng(seed).integers(0, 2, (h, w)) * 255).astype(np.uint8)
    if kind == "dot_lattice":
        img = np.zeros((h, w), np.uint8)
        ii, jj = np.meshgrid(np.arange(0, w, 2) // 2, np.arange(0, h, 2) // 2)
        img[::2, ::2] = (50 + (ii * 24 + jj) % 192).astype(np.uint8)
        return img
    if kind == "low_contrast":
        rng = np.random.default_rng(seed)
        base = rng.integers(100, 119, (h // 3 + 1, w // 3 + 1))
        return np.kron(base, np.ones((3, 3), np.int64))[:h, :w].astype(np.uint8)
    return np.full((h, w), 117, np.uint8)


@pytest.mark.parametrize("name,w,h,nf,seed,kind", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_stage_and_final_parity(pkg, name, w, h, nf, seed, kind):
    img = make_image(w, h, seed, kind)
    params = (nf, 1.2, 8, 20, 7)
    o = oracle.OracleExtractor(*params)
    ok, od = o(img)
    g = pkg.ORBextractor(*params)
    gk, gd = g(img)
    assert np.array_equal(g.GetScaleFactors(), o.scale) and np.array_equal(g.GetInverseScaleFactors(), o.inv_scale)
    assert np.array_equal(g.GetScaleSigmaSquares(), o.sigma2) and np.array_equal(g.GetInverseScaleSigmaSquares(), o.inv_sigma2)
    assert np.array_equal(g.mnFeaturesPerLevel, o.quota)
    pyr = g.mvImagePyramid
    for l in range(8):
        assert g.level_size(l) == o.level_size(l)
        assert np.array_equal(pyr[l], o.level(l, 0)), "mvImagePyramid[%d]" % l
        assert np.array_equal(g.pyramid_level(l, with_border=True), o.level(l, 1)), "border level %d" % l
        assert np.array_equal(g.stage(l, 0), o.stage(l, 0)), "FAST candidates level %d" % l
        assert np.array_equal(g.stage(l, 1), o.stage(l, 1)), "octree selection level %d" % l
        ob = o.level(l, 2)
        if ob is not None:
            assert np.array_equal(g.blurred_level(l), ob), "blur level %d" % l
    assert_same_keypoints(gk, gd, ok, od)
    if kind == "flat":
        assert len(gk) == 0 and gd.shape == (0, 32)


def test_kitti_thresholds_variant(pkg):
    """KITTI04-12.yaml uses iniThFAST = 12 (Examples/Stereo/KITTI04-12.yaml:50)."""
    img = gen_synth.texture(376, 1241, 12)
    params = (2000, 1.2, 8, 12, 7)
    ok, od = oracle.OracleExtractor(*params)(img)
    gk, gd = pkg.ORBextractor(*params)(img)
    assert_same_keypoints(gk, gd, ok, od)


def test_other_pyramid_parameters(pkg):
    img = gen_synth.texture(480, 640, 21)
    # 1 and 2 levels, a 1.05 pyramid, iniThFAST == minThFAST (the fallback never adds anything), thresholds 40 / 39
    # (a handful of keypoints: iniS = 1 in the kernel's score units)
    for params in [(800, 1.5, 4, 25, 10), (1500, 1.1, 12, 20, 7), (300, 2.0, 3, 30, 5), (500, 1.2, 1, 20, 7), (700, 1.2, 2, 20, 7),
                   (1000, 1.05, 8, 20, 7), (600, 1.2, 8, 7, 7), (600, 1.2, 8, 40, 39)]:
        ok, od = oracle.OracleExtractor(*params)(img)
        gk, gd = pkg.ORBextractor(*params)(img)
        assert_same_keypoints(gk, gd, ok, od)


def test_batch_equals_single_calls(pkg):
    imgs = gen_synth.frame_batch(6, 376, 1241, 77)
    params = (2000, 1.2, 8, 20, 7)
    g = pkg.ORBextractor(*params, max_width=1241, max_height=376, max_batch=6)
    kps, desc, counts = g.extract_batch(imgs)
    o = oracle.OracleExtractor(*params)
    for i in range(6):
        ok, od = o(imgs[i])
        assert counts[i] == len(ok)
        assert_same_keypoints(kps[i, :counts[i]], desc[i, :counts[i]], ok, od)
    # mvImagePyramid of a non-zero batch slot
    assert np.array_equal(g.pyramid_level(3, image=5), o.level(3, 0))


def test_chunked_host_pipeline_equals_single_calls(pkg):
    """Batches of >= 32 images go through the 4-chunk H2D / kernels / D2H pipeline of orbgpu_extract_batch."""
    imgs = gen_synth.frame_batch(37, 240, 416, 5)
    params = (500, 1.2, 8, 20, 7)
    g = pkg.ORBextractor(*params, max_width=416, max_height=240, max_batch=37)
    kps, desc, counts = g.extract_batch(imgs)
    o = oracle.OracleExtractor(*params)
    for i in range(37):
        ok, od = o(imgs[i])
        assert counts[i] == len(ok), i
        assert_same_keypoints(kps[i, :counts[i]], desc[i, :counts[i]], ok, od)
    # the device-resident results of the same batch are what the host received
    k2, d2, c2 = g.download_results()
    assert np.array_equal(c2, counts) and np.array_equal(d2[5, :c2[5]], desc[5, :counts[5]])


def test_stage_timers_and_pipe_probes(pkg):
    import ctypes
    g = pkg.ORBextractor(500, 1.2, 8, 20, 7, max_width=416, max_height=240, max_batch=4)
    g.set_profiling(True)
    g.extract_batch(gen_synth.frame_batch(4, 240, 416, 6))
    t = g.stage_times()
    assert set(t) == {"pyramid", "fast", "octree", "blur", "describe", "stereo"}
    assert all(t[k] > 0 for k in ("pyramid", "fast", "octree", "blur", "describe")) and t["stereo"] == 0
    popc, lop3 = ctypes.c_double(), ctypes.c_double()
    pkg._lib.check(pkg.lib().orbgpu_measure_popc_peak(-1, ctypes.byref(popc), ctypes.byref(lop3)))
    assert 1e12 < popc.value < 1e13 and 5e12 < lop3.value < 5e13      # 16 and 64 lanes/clk/SM on B200
    assert pkg.lib().orbgpu_launch_count(0) > 0


def test_handle_reuse_across_sizes_and_row_stride(pkg):
    params = (1000, 1.2, 8, 20, 7)
    g = pkg.ORBextractor(*params, max_width=800, max_height=600)
    o = oracle.OracleExtractor(*params)
    big = gen_synth.texture(600, 800, 31)
    for view in (big, big[:480, :640], big[10:400, 5:700], big):      # non-contiguous rows: step > width
        ok, od = o(np.ascontiguousarray(view))
        gk, gd = g(view)
        assert_same_keypoints(gk, gd, ok, od)


def test_4k_8000_features(pkg):
    """BASELINE.json configs[4]: 3840 x 2160, nFeatures = 8000."""
    img = gen_synth.frame_batch(1, 2160, 3840, 4000)[0]
    params = (8000, 1.2, 8, 20, 7)
    o = oracle.OracleExtractor(*params)
    ok, od = o(img)
    g = pkg.ORBextractor(*params)
    gk, gd = g(img)
    for l in range(8):
        assert np.array_equal(g.stage(l, 1), o.stage(l, 1)), "octree selection level %d" % l
    assert_same_keypoints(gk, gd, ok, od)
    assert len(gk) >= 8000


def test_empty_and_unsupported_inputs(pkg):
    g = pkg.ORBextractor(500, 1.2, 8, 20, 7)
    k, d = g(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)                       # ORBextractor.cc:981-982
    with pytest.raises(pkg.OrbGpuError) as e:                       # top level smaller than one FAST cell
        g(np.zeros((120, 160), np.uint8))
    assert e.value.code == -4
    with pytest.raises(pkg.OrbGpuError) as e:                       # aspect < 0.5: nIni == 0 in the reference
        g(gen_synth.texture(900, 300, 1))
    assert e.value.code == -4
    with pytest.raises(ValueError):
        g(np.zeros((100, 100, 3), np.uint8))


def test_two_handles_from_two_threads(pkg):
    """The stereo constructor runs the left and right extractor concurrently (Frame.cc:87-90)."""
    import threading
    left, right = gen_synth.stereo_pair(376, 1241, 3)
    params = (2000, 1.2, 8, 20, 7)
    gl, gr = pkg.ORBextractor(*params, max_width=1241, max_height=376), pkg.ORBextractor(*params, max_width=1241, max_height=376)
    out = {}
    ts = [threading.Thread(target=lambda k, g, im: out.__setitem__(k, g(im)), args=a) for a in (("l", gl, left), ("r", gr, right))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    o = oracle.OracleExtractor(*params)
    assert_same_keypoints(*out["l"], *o(left))
    assert_same_keypoints(*out["r"], *o(right))


def test_golden_fixture(pkg):
    """tests/golden/orb_320x240.npz was produced by the cv2-driven restatement (real OpenCV
    primitives) with tests/golden/make_golden.py."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "orb_320x240.npz"))
    nf, sf, nl, ini, mn = z["params"]
    gk, gd = pkg.ORBextractor(int(nf), float(sf), int(nl), int(ini), int(mn))(z["image"])
    assert np.array_equal(gd, z["descriptors"])
    for f in gk.dtype.names:
        assert np.array_equal(gk[f], z["keypoints"][f]), f


def test_two_graphed_chunks_and_repeat_calls(pkg):
    """A 16-image host batch is two chunks of 8, both on the captured-graph path (one graph slot per chunk, first call on a
    fresh handle: the descriptor pattern table must already be on the device, outside the capture).  Repeated calls replay."""
    import ctypes as C
    imgs = gen_synth.frame_batch(16, 240, 416, 9)
    params = (500, 1.2, 8, 20, 7)
    g = pkg.ORBextractor(*params, max_width=416, max_height=240, max_batch=16)
    o = oracle.OracleExtractor(*params)
    want = [o(imgs[i]) for i in range(16)]
    for _ in range(3):
        kps, desc, counts = g.extract_batch(imgs)
        for i in range(16):
            assert counts[i] == len(want[i][0]), i
            assert_same_keypoints(kps[i, :counts[i]], desc[i, :counts[i]], *want[i])
    assert pkg.lib().orbgpu_last_error() in (b"", None) or b"capture" not in pkg.lib().orbgpu_last_error()


def test_one_pending_batch_per_handle(pkg):
    """A second _async (or a device-resident extraction) before orbgpu_wait would overwrite results under the first batch's
    D2H copies: it is rejected, and the handle works again after the wait."""
    import ctypes as C
    from refactored_orb_slam2_b200._lib import KP_DTYPE, E_INVALID, ptr
    imgs = gen_synth.frame_batch(4, 240, 416, 3)
    params = (500, 1.2, 8, 20, 7)
    g = pkg.ORBextractor(*params, max_width=416, max_height=240, max_batch=4)
    cap = g.capacity
    kps, desc, counts = np.empty((4, cap), KP_DTYPE), np.empty((4, cap, 32), np.uint8), np.zeros(4, np.int32)
    L = pkg.lib()
    args = (g.handle, ptr(imgs), 4, 416, 240, 416, 240 * 416, ptr(kps), ptr(desc), cap, ptr(counts))
    assert L.orbgpu_extract_batch_async(*args) == 0
    assert L.orbgpu_extract_batch_async(*args) == E_INVALID
    assert L.orbgpu_extract_batch_device(g.handle, ptr(imgs), 4, 416, 240, 416, 240 * 416) == E_INVALID
    assert L.orbgpu_wait(g.handle) == 0
    ok, od = oracle.OracleExtractor(*params)(imgs[2])
    assert_same_keypoints(kps[2, :counts[2]], desc[2, :counts[2]], ok, od)
    assert L.orbgpu_extract_batch_async(*args) == 0 and L.orbgpu_wait(g.handle) == 0
