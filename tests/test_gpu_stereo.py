"""Frame::ComputeStereoMatches parity: mvuRight / mvDepth bit-exact against the oracle (north_star
allows 1e-3 px; the float sequence is reproduced exactly, so the tolerance is 0)."""
import numpy as np
import pytest

import gen_synth
import oracle

pytestmark = pytest.mark.gpu
KITTI = (386.1448, 718.856)      # bf, fx: Examples/Stereo/KITTI00-02.yaml:8,25
EUROC = (47.90639384423901, 435.2046959714599)   # Examples/Stereo/EuRoC.yaml:8,25


@pytest.fixture(scope="module")
def pkg(built):
    import refactored_orb_slam2_b200 as p
    return p


def oracle_pair(params, left, right, bf, fx):
    ol, orr = oracle.OracleExtractor(*params), oracle.OracleExtractor(*params)
    kl, dl = ol(left)
    kr, dr = orr(right)
    u, z, sad, kept = oracle.stereo(ol, orr, kl, dl, kr, dr, bf, np.float32(np.float32(bf) / np.float32(fx)))
    return len(kl), u, z, kept


@pytest.mark.parametrize("w,h,nf,cam,seed", [(1241, 376, 2000, KITTI, 2), (752, 480, 1200, EUROC, 1000), (640, 480, 1000, KITTI, 5)])
def test_stereo_pair_parity(pkg, w, h, nf, cam, seed):
    left, right = gen_synth.stereo_pair(h, w, seed)
    params = (nf, 1.2, 8, 20, 7)
    bf, fx = cam
    n, ou, oz, kept = oracle_pair(params, left, right, bf, fx)
    f = pkg.Frame(left, right, pkg.ORBextractor(*params), pkg.ORBextractor(*params), bf, fx)
    assert f.N == n
    assert np.array_equal(f.mvuRight, ou)
    assert np.array_equal(f.mvDepth, oz)
    assert kept > n // 4, "the synthetic pair should yield plenty of stereo matches"
    assert int((f.mvuRight >= 0).sum()) == kept


@pytest.mark.parametrize("npairs", [5, 11])        # <= 8 pairs run the latency thread shapes of octree / describe / stereo, more the throughput ones
def test_batched_pairs_one_handle(pkg, npairs):
    """EuRoC-shaped pairs batched on one handle: left images first, right images after them."""
    left, right = gen_synth.frame_batch(npairs, 480, 752, 1000, stereo=True)
    params = (1200, 1.2, 8, 20, 7)
    bf, fx = EUROC
    g = pkg.ORBextractor(*params, max_width=752, max_height=480, max_batch=2 * npairs)
    kps, desc, counts = g.extract_batch(np.concatenate([left, right]))
    u, z = pkg.ComputeStereoMatches(g, g, bf, np.float32(bf) / np.float32(fx), first_left=0, first_right=npairs, pairs=npairs)
    for i in range(npairs):
        n, ou, oz, kept = oracle_pair(params, left[i], right[i], bf, fx)
        assert counts[i] == n
        assert np.array_equal(u[i, :n], ou) and np.array_equal(z[i, :n], oz), "pair %d" % i


def test_stereo_edge_cases_in_one_batch(pkg):
    """Pairs with no right keypoints, with no left keypoints, a 2-px dot lattice (crowded row buckets, many equal descriptors
    distances) and 0 / 255 noise, batched on one handle: every pair equals the oracle and the empty ones do not disturb the rest."""
    h, w = 263, 341
    def lattice(width):
        img = np.zeros((h, width), np.uint8)
        ii, jj = np.meshgrid(np.arange(0, width, 2) // 2, np.arange(0, h, 2) // 2)
        img[::2, ::2] = (50 + (ii * 24 + jj) % 192).astype(np.uint8)
        return img
    tex, flat = gen_synth.texture(h, w, 3), np.full((h, w), 90, np.uint8)
    lat = lattice(w + 6)
    noise = (np.random.default_rng(4).integers(0, 2, (h, w + 10)) * 255).astype(np.uint8)
    pairs = [(tex, flat), (flat, tex), (lat[:, :-6], lat[:, 6:]), (noise[:, :-10], noise[:, 10:])]
    left = np.stack([np.ascontiguousarray(p[0]) for p in pairs])
    right = np.stack([np.ascontiguousarray(p[1]) for p in pairs])
    params = (500, 1.2, 8, 20, 7)
    bf, fx = KITTI
    g = pkg.ORBextractor(*params, max_width=w, max_height=h, max_batch=2 * len(pairs))
    kps, desc, counts = g.extract_batch(np.concatenate([left, right]))
    u, z = pkg.ComputeStereoMatches(g, g, bf, np.float32(bf) / np.float32(fx), first_left=0, first_right=len(pairs), pairs=len(pairs))
    matched = []
    for i in range(len(pairs)):
        n, ou, oz, kept = oracle_pair(params, left[i], right[i], bf, fx)
        assert counts[i] == n
        assert np.array_equal(u[i, :n], ou) and np.array_equal(z[i, :n], oz), "pair %d" % i
        matched.append(kept)
    assert matched[0] == 0 and matched[1] == 0 and matched[2] > 100 and matched[3] > 40
    # the reference's constructor shape (one pair per call) on the empty cases
    for l, r in pairs[:2]:
        f = pkg.Frame(l, r, pkg.ORBextractor(*params), pkg.ORBextractor(*params), bf, fx)
        assert (f.mvuRight == -1).all() and (f.mvDepth == -1).all() and len(f.mvuRight) == f.N


def test_integer_shift_recovers_disparity(pkg):
    """Size-independent property: right = left shifted by d px  =>  uL - uR ~= d for the matches."""
    d = 23
    canvas = gen_synth.texture(376, 1241 + d, 9)
    left, right = np.ascontiguousarray(canvas[:, :-d]), np.ascontiguousarray(canvas[:, d:])   # uR = uL - d
    f = pkg.Frame(left, right, pkg.ORBextractor(2000, 1.2, 8, 20, 7), pkg.ORBextractor(2000, 1.2, 8, 20, 7), *KITTI)
    m = f.mvuRight >= 0
    assert m.sum() > 500
    disp = f.mvKeys["x"][m] - f.mvuRight[m]
    assert abs(np.median(disp) - d) < 0.5
    assert np.allclose(f.mvDepth[m], np.float32(KITTI[0]) / disp, rtol=1e-6)


def test_pipeline_equals_sequential_calls(pkg):
    """StereoBatchPipeline (two handles, async entry points + orbgpu_wait, batches in flight) returns exactly what sequential
    extract_batch + ComputeStereoMatches calls return."""
    npairs, params = 3, (1200, 1.2, 8, 20, 7)
    bf, fx = EUROC
    mb = float(np.float32(bf) / np.float32(fx))
    batches = []
    for seed in (1, 2, 3, 4, 5):
        l, r = gen_synth.frame_batch(npairs, 480, 752, seed, stereo=True)
        batches.append(np.ascontiguousarray(np.concatenate([l, r])))
    pipe = pkg.StereoBatchPipeline(params, 752, 480, npairs, bf, mb, depth=2, pin_outputs=False)
    g = pkg.ORBextractor(*params, max_width=752, max_height=480, max_batch=2 * npairs)
    got = []
    for i, b in enumerate(batches):
        pipe.submit(b)
        if i >= 1:
            k, d, c, u, z = pipe.result()
            got.append((k.copy(), d.copy(), c.copy(), u.copy(), z.copy()))
    k, d, c, u, z = pipe.result()
    got.append((k.copy(), d.copy(), c.copy(), u.copy(), z.copy()))
    pipe.close()
    for b, (k, d, c, u, z) in zip(batches, got):
        k2, d2, c2 = g.extract_batch(b)
        u2, z2 = pkg.ComputeStereoMatches(g, g, bf, mb, 0, npairs, npairs)
        assert np.array_equal(c, c2)
        for i in range(2 * npairs):
            assert np.array_equal(d[i, :c[i]], d2[i, :c2[i]])
            assert np.array_equal(k[i, :c[i]]["angle"], k2[i, :c2[i]]["angle"])
        for i in range(npairs):
            assert np.array_equal(u[i, :c[i]], u2[i, :c2[i]]) and np.array_equal(z[i, :c[i]], z2[i, :c2[i]])


def test_left_and_right_handles_must_share_nfeatures(pkg):
    """The row buckets are sized by the left handle's keypoint capacity: a right extractor with more features is rejected
    instead of writing past them."""
    from refactored_orb_slam2_b200._lib import OrbGpuError
    h, w = 240, 416
    left, right = gen_synth.stereo_pair(h, w, 5)
    gl = pkg.ORBextractor(300, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    gr = pkg.ORBextractor(900, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    gl.extract_batch(left[None])
    gr.extract_batch(right[None])
    bf, fx = KITTI
    with pytest.raises(OrbGpuError):
        pkg.ComputeStereoMatches(gl, gr, bf, np.float32(bf) / np.float32(fx), first_left=0, first_right=0, pairs=1)
