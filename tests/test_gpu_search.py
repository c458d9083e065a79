"""GPU parity of the frame grid and the ORBmatcher candidate loops (csrc/search.cu) through the C ABI.

Three layers: (1) grid / GetFeaturesInArea against the oracle; (2) every search policy — window and list sources,
level / stereo / chi-square / epipolar gates, both ratio tests, both greedy modes — record by record against
oracle/oracle_search.cpp on scenes with heavy keypoint contention; (3) shim/ORBmatcher.cc on liborbgpu.so against the
reference's own ORBmatcher.cc (oracle/_ref, built in the development container) on the eleven searches.  Bit-exact."""
import numpy as np
import pytest

import gen_synth
import oracle
import slam_ref
import slam_scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import refactored_orb_slam2_b200 as p
    return p


def make_view(seed, npoints=1400, clutter=500):
    W = slam_ref.World(seed, npoints=npoints, clutter=clutter)
    return W, W.observe(slam_ref._pose(0.004, -0.02, 0.002, [-0.55, 0.02, -0.45]))


def both(pkg, V, keyframe=False, cam=slam_ref.KITTI):
    geo = pkg.grid_geometry(0.0, float(cam["w"]), 0.0, float(cam["h"]), keyframe=keyframe)
    return pkg.FrameFeatures(V.kp, V.desc, V.uR, geo), oracle.OracleFeatures(V.kp, V.desc, V.uR, geo)


def random_window_queries(pkg, rng, V, n, radius=(2, 60)):
    from refactored_orb_slam2_b200._lib import QUERY_DTYPE
    q = np.zeros(n, QUERY_DTYPE)
    src = rng.integers(0, len(V.kp), n)
    q["u"] = V.kp["x"][src] + rng.normal(0, 3, n)
    q["v"] = V.kp["y"][src] + rng.normal(0, 3, n)
    q["r"] = rng.uniform(*radius, n)
    lv = V.kp["octave"][src]
    q["min_level"] = lv - rng.integers(0, 2, n)
    q["max_level"] = np.where(rng.random(n) < 0.15, -1, lv + rng.integers(0, 2, n))
    q["ur"] = q["u"] - rng.uniform(2, 60, n)
    q["ur_tol"] = q["r"]
    q["flags"] = np.where(rng.random(n) < 0.85, 1, 0) | np.where(rng.random(n) < 0.5, 2, 0)
    desc = V.desc[src].copy()
    flip = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
    desc ^= np.where(rng.random((n, 1)) < 0.7, flip & rng.integers(0, 256, (n, 32), dtype=np.uint8), 0).astype(np.uint8)
    return q, desc, src


def test_grid_and_features_in_area(pkg):
    for seed, kf in ((1, False), (2, True)):
        W, V = make_view(seed)
        V.kp["x"][:5] = [-3.0, 1240.9, 1241.0, 640.5, 0.0]        # outside / on the edge of the grid
        V.kp["y"][:5] = [10.0, 375.9, 376.0, -0.2, 0.0]
        g, o = both(pkg, V, kf)
        gs, gi = g.grid()
        os_, oi = o.grid()
        assert np.array_equal(gs, os_) and np.array_equal(gi, oi)
        rng = np.random.default_rng(seed)
        q, _, _ = random_window_queries(pkg, rng, V, 600, radius=(0.5, 120))
        q["u"][:20] = rng.uniform(-200, 1500, 20)                   # windows partly / fully outside the image
        q["v"][:20] = rng.uniform(-200, 600, 20)
        go, gitems = g.GetFeaturesInArea(q["u"], q["v"], q["r"], q["min_level"], q["max_level"])
        oo, oitems = o.in_area(q)
        assert np.array_equal(go, oo) and np.array_equal(gitems, oitems)
        assert go[-1] > 1000
        k, u, d = g.download()
        assert np.array_equal(k, V.kp) and np.array_equal(u, V.uR) and np.array_equal(d, V.desc)


POLICIES = [
    ("local_points", dict(ratio_mode=1, greedy=1, accept_max=100, ratio=0.8, stereo_gate=1)),
    ("last_frame", dict(greedy=1, accept_max=100, stereo_gate=1)),
    ("reloc", dict(greedy=1, accept_max=64)),
    ("kf_projection", dict(greedy=1, accept_max=50)),
    ("initialization", dict(ratio_mode=2, greedy=2, accept_max=50, ratio=0.9)),
    ("fuse", dict(gate=1, accept_max=50)),
    ("sim3", dict(accept_max=100)),
]


@pytest.mark.parametrize("name,kw", POLICIES, ids=[p[0] for p in POLICIES])
def test_window_search_policies_vs_oracle(pkg, name, kw):
    W, V = make_view(7)
    g, o = both(pkg, V)
    rng = np.random.default_rng(11)
    nq = 3000                                    # 1.6 queries per keypoint: heavy contention for the greedy modes
    q, desc, _ = random_window_queries(pkg, rng, V, nq)
    occ = (rng.random(len(V.kp)) < 0.15).astype(np.uint8)
    sig = np.float32(1.2) ** (2 * np.arange(16, dtype=np.float32))
    P = pkg.make_policy(source=0, inv_level_sigma2=(1 / sig).astype(np.float32), **kw)
    got, passes = g.search(q, desc, P, occupied=occ)
    want = o.search(q, desc, P, occupied=occ)
    assert np.array_equal(got, want)
    assert want["accepted"].sum() > 200
    if kw.get("greedy"):
        assert passes >= 3, "the scene should force the greedy fixed-point iteration to do real work"
        free = pkg.make_policy(source=0, inv_level_sigma2=(1 / sig).astype(np.float32), **dict(kw, greedy=0))
        assert not np.array_equal(o.search(q, desc, free, occupied=occ), want)
    else:
        assert passes == 1


@pytest.mark.parametrize("mode", ["bow_frame", "bow_kf", "triangulation", "triangulation_stereo"])
def test_list_search_policies_vs_oracle(pkg, mode):
    from refactored_orb_slam2_b200._lib import QUERY_DTYPE
    W, V = make_view(9)
    g, o = both(pkg, V, keyframe=True)
    rng = np.random.default_rng(5)
    n = len(V.kp)
    order = np.argsort(V.node, kind="stable").astype(np.int32)          # one candidate list per vocabulary node
    nodes, counts = np.unique(V.node, return_counts=True)
    start = np.concatenate([[0], np.cumsum(counts)])
    nq = 2500
    qnode = rng.integers(0, len(nodes), nq)
    q = np.zeros(nq, QUERY_DTYPE)
    q["min_level"], q["max_level"] = -1, -1
    q["list_begin"], q["list_end"] = start[qnode], start[qnode + 1]
    src = order[(start[qnode] + rng.integers(0, 1 << 30, nq) % counts[qnode])]
    q["u"], q["v"] = V.kp["x"][src] + rng.normal(0, 2, nq), V.kp["y"][src] + rng.normal(0, 2, nq)
    q["flags"] = 1 | np.where(rng.random(nq) < 0.5, 2, 0)
    desc = V.desc[src].copy()
    desc[:, :2] ^= rng.integers(0, 256, (nq, 2), dtype=np.uint8)
    occ = (rng.random(n) < 0.2).astype(np.uint8)
    scale = (np.float32(1.2) ** np.arange(16, dtype=np.float32)).astype(np.float32)
    if mode == "bow_frame":
        P = pkg.make_policy(source=1, ratio_mode=2, greedy=1, accept_max=50, ratio=0.7)
    elif mode == "bow_kf":
        P = pkg.make_policy(source=1, ratio_mode=2, greedy=1, accept_max=49, ratio=0.75)
    else:
        F12 = slam_scenarios._f12(slam_ref._pose(0, 0, 0, [0, 0, 0]), slam_ref._pose(0.0, -0.002, 0.0, [-0.05, 0.0, -0.04]), slam_ref.KITTI)
        P = pkg.make_policy(source=1, gate=2, accept_max=50, only_stereo=int(mode.endswith("stereo")), ex=600.0, ey=180.0, f12=F12,
                            level_sigma2=scale * scale, scale_factors=scale)
    got, passes = g.search(q, desc, P, occupied=occ, cand=order)
    want = o.search(q, desc, P, occupied=occ, cand=order)
    assert np.array_equal(got, want)
    assert want["accepted"].sum() > 100
    assert (passes >= 3) if mode.startswith("bow") else (passes == 1)


def test_search_edge_cases(pkg):
    from refactored_orb_slam2_b200._lib import KP_DTYPE, QUERY_DTYPE
    geo = pkg.grid_geometry(0.0, 640.0, 0.0, 480.0)
    empty = pkg.FrameFeatures(np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), None, geo)
    q = np.zeros(3, QUERY_DTYPE)
    q["u"], q["v"], q["r"] = 100, 100, 50
    m, _ = empty.search(q, np.zeros((3, 32), np.uint8), pkg.make_policy())
    assert (m["idx"] == -1).all() and (m["dist"] == 256).all() and (m["accepted"] == 0).all()
    W, V = make_view(3, 300, 50)
    g, o = both(pkg, V)
    m, passes = g.search(np.zeros(0, QUERY_DTYPE), np.zeros((0, 32), np.uint8), pkg.make_policy(greedy=1))
    assert len(m) == 0
    # monocular set (no right coordinates) and identical descriptors everywhere: ties go to the first candidate in walk order
    same = np.tile(V.desc[:1], (len(V.kp), 1))
    geo2 = pkg.grid_geometry(0.0, 1241.0, 0.0, 376.0)
    gm, om = pkg.FrameFeatures(V.kp, same, None, geo2), oracle.OracleFeatures(V.kp, same, None, geo2)
    rng = np.random.default_rng(0)
    q, _, _ = random_window_queries(pkg, rng, V, 400, radius=(20, 90))
    d = np.tile(V.desc[:1], (400, 1))
    for kw in (dict(greedy=1, accept_max=100), dict(ratio_mode=2, ratio=0.9, accept_max=50, greedy=2), dict(ratio_mode=1, ratio=0.8, stereo_gate=1)):
        P = pkg.make_policy(**kw)
        got, _ = gm.search(q, d, P)
        assert np.array_equal(got, om.search(q, d, P))
    assert (gm.download()[1] == -1).all()
    with pytest.raises(pkg.OrbGpuError):
        bad = q.copy(); bad["list_end"] = 5
        g.search(bad, d, pkg.make_policy(source=1), cand=np.array([0, 1, 99999], np.int32))


def test_features_from_extractor_equal_host_built_set(pkg):
    left, right = gen_synth.stereo_pair(376, 1241, 4)
    params = (2000, 1.2, 8, 20, 7)
    ex = pkg.ORBextractor(*params, max_width=1241, max_height=376, max_batch=2)
    kps, desc, counts = ex.extract_batch(np.stack([left, right]))
    mbf, fx = 386.1448, 718.856
    u, z = pkg.ComputeStereoMatches(ex, ex, mbf, mbf / fx, first_left=0, first_right=1, pairs=1)
    n = counts[0]
    geo = pkg.grid_geometry(0.0, 1241.0, 0.0, 376.0)
    a = pkg.FrameFeatures.from_extractor(ex, 0, geo, pair=0)
    b = pkg.FrameFeatures(kps[0, :n], desc[0, :n], u[0, :n], geo)
    o = oracle.OracleFeatures(kps[0, :n], desc[0, :n], u[0, :n], geo)
    assert a.n == b.n == n
    for x, y in zip(a.grid(), b.grid()):
        assert np.array_equal(x, y)
    for x, y in zip(a.grid(), o.grid()):
        assert np.array_equal(x, y)
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x, y)
    mono = pkg.FrameFeatures.from_extractor(ex, 1, geo)
    assert mono.n == counts[1] and (mono.download()[1] == -1).all()
    # self-search: every keypoint finds itself at distance 0 inside a 1-pixel window
    from refactored_orb_slam2_b200._lib import QUERY_DTYPE
    q = np.zeros(n, QUERY_DTYPE)
    q["u"], q["v"], q["r"], q["min_level"], q["max_level"] = kps[0, :n]["x"], kps[0, :n]["y"], 1.0, kps[0, :n]["octave"], kps[0, :n]["octave"]
    m, _ = a.search(q, desc[0, :n], pkg.make_policy(accept_max=0))
    assert (m["dist"] == 0).all() and m["accepted"].all()
    assert np.array_equal(m, o.search(q, desc[0, :n], pkg.make_policy(accept_max=0)))


@pytest.mark.skipif(not (slam_ref.available("ref") and slam_ref.available("shim_gpu")), reason="oracle/_ref checker libraries not shipped with this snapshot")
@pytest.mark.parametrize("seed,small", [(3, True), (5, False), (11, False)])
def test_shim_matcher_on_gpu_equals_reference_matcher(pkg, seed, small):
    """shim/ORBmatcher.cc (host policy) + liborbgpu.so (candidate loops) against the reference's OWN ORBmatcher.cc on the same
    pointer-graph scenes: every output of the eleven searches, driven as their callers drive them."""
    ref = slam_scenarios.run_all("ref", seed, small)
    launches0 = pkg.lib().orbgpu_launch_count(0)
    mine = slam_scenarios.run_all("shim_gpu", seed, small)
    assert slam_scenarios.compare(ref, mine) == []
    assert len(ref) == 21


@pytest.mark.skipif(not slam_ref.available("shim_gpu"), reason="oracle/_ref checker libraries not shipped with this snapshot")
@pytest.mark.parametrize("w,h,nf,seed,cam", [(1241, 376, 2000, 2, (718.856, 607.1928, 185.2157, 386.1448)), (752, 480, 1200, 1000, (435.2047, 367.4517, 252.2008, 47.9064))])
def test_reference_frame_constructor_linked_against_the_drop_in(pkg, w, h, nf, seed, cam):
    """The reference's OWN Frame::Frame(stereo) (Frame.cc:66-127, unmodified) compiled against shim/ORBextractor + shim/ORBmatcher
    on liborbgpu.so: keypoints, descriptors, mvuRight / mvDepth (the reference's host ComputeStereoMatches on the pyramid the shim
    downloads) and the grid must equal the oracle's; so must the device replacement of the method body (shim/FrameStereo.h)."""
    fx, cx, cy, bf = cam
    left, right = gen_synth.stereo_pair(h, w, seed)
    params = (nf, 1.2, 8, 20, 7)
    out = slam_ref.SlamLib("shim_gpu").stereo_frame(params, left, right, fx, fx, cx, cy, bf)
    ol, orr = oracle.OracleExtractor(*params), oracle.OracleExtractor(*params)
    (kl, dl), (kr, dr) = ol(left), orr(right)
    assert len(out["keys"]) == len(kl)
    for f in kl.dtype.names:
        assert np.array_equal(out["keys"][f], kl[f]), f
    assert np.array_equal(out["desc"], dl)
    mb = np.float32(np.float32(bf) / np.float32(fx))
    ou, oz, _, _ = oracle.stereo(ol, orr, kl, dl, kr, dr, bf, mb)
    assert np.array_equal(out["u_right"], ou) and np.array_equal(out["depth"], oz)
    assert np.array_equal(out["u_right_device"], ou) and np.array_equal(out["depth_device"], oz)
    geo = pkg.grid_geometry(0.0, float(w), 0.0, float(h))
    start, items = oracle.OracleFeatures(kl, dl, ou, geo).grid()
    assert np.array_equal(start, out["cell_start"]) and np.array_equal(items, out["cell_items"])


def test_concurrent_searches_from_three_threads_share_one_feature_set(pkg):
    """Tracking, LocalMapping and LoopClosing search concurrently in the reference (SURVEY 8b threading): per-thread stream and
    scratch inside orbgpu_search, read-only feature sets."""
    import threading
    W, V = make_view(13)
    g, o = both(pkg, V)
    jobs = []
    for t in range(3):
        rng = np.random.default_rng(100 + t)
        q, d, _ = random_window_queries(pkg, rng, V, 1500 + 400 * t)
        P = pkg.make_policy(greedy=1 if t != 1 else 2, ratio_mode=2 if t == 1 else 0, ratio=0.9, accept_max=100, stereo_gate=int(t == 0))
        jobs.append((q, d, P, o.search(q, d, P)))
    bad = []

    def work(job):
        q, d, P, want = job
        for _ in range(25):
            got, _ = g.search(q, d, P)
            if not np.array_equal(got, want):
                bad.append(1)

    th = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not bad


def test_distinctive_descriptors_vs_oracle(pkg):
    """Row f4: batched MapPoint::ComputeDistinctiveDescriptors — groups of 1..70 observations, ties, empty groups."""
    rng = np.random.default_rng(4)
    sizes = np.concatenate([rng.integers(1, 12, 3000), rng.integers(12, 70, 200), [0, 1, 2, 2, 33, 64, 65]])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    rows = np.repeat(base, sizes, axis=0)
    flip = rng.integers(0, 256, rows.shape, dtype=np.uint8) & rng.integers(0, 256, rows.shape, dtype=np.uint8) & rng.integers(0, 256, rows.shape, dtype=np.uint8)
    rows ^= flip
    rows[offsets[-3]:offsets[-3] + 10] = rows[offsets[-3]]          # identical rows: the first wins
    gi, gm = pkg.distinctive_descriptors(rows, offsets)
    oi, om = oracle.distinctive(rows, offsets)
    assert np.array_equal(gi, oi) and np.array_equal(gm, om)
    assert gi[len(sizes) - 7] == -1
    i0, m0 = pkg.distinctive_descriptors(np.zeros((0, 32), np.uint8), np.zeros(1, np.int32))
    assert len(i0) == 0
