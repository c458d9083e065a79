"""Pins the ORBmatcher row and Frame::ComputeStereoMatches to the reference's OWN code (CPU only).

oracle/_ref/libref_slam.so is the reference's unmodified Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc, ORBmatcher.cc and
ORBextractor.cc compiled from /root/reference against the OpenCV API stand-in of oracle/ref_shim.  Against it:

  * every one of the eleven searches, driven as its caller drives it on a synthetic map, must leave exactly the same
    state behind when shim/ORBmatcher.cc (the product's host policy) runs on the CPU oracle backend
    (oracle/_ref/libshim_slam_cpu.so) — this pins the host policy AND oracle/oracle_search.cpp's candidate loops;
  * the oracle's restatement of Frame::ComputeStereoMatches must give bit-identical mvuRight / mvDepth;
  * ORBmatcher::DescriptorDistance must agree with the oracle and with numpy popcount.
Skipped where the checker can be neither found nor built (the GPU box only has the prebuilt files)."""
import numpy as np
import pytest

import gen_synth
import oracle
import slam_ref
import slam_scenarios

pytestmark = pytest.mark.skipif(not (slam_ref.available("ref") and slam_ref.available("shim_cpu")),
                                reason="oracle/_ref/libref_slam.so absent and /root/reference not present")


@pytest.mark.parametrize("seed,small", [(3, True), (5, False), (11, False)])
def test_shim_host_policy_on_oracle_backend_equals_reference_matcher(seed, small):
    ref = slam_scenarios.run_all("ref", seed, small)
    mine = slam_scenarios.run_all("shim_cpu", seed, small)
    assert set(ref) == set(mine) and len(ref) == 21
    assert slam_scenarios.compare(ref, mine) == []
    # the scenes are not degenerate: every search finds matches
    for name, outs in ref.items():
        if "n" in outs:
            assert int(outs["n"]) > 10, name


def test_reference_matcher_is_deterministic_on_these_scenes():
    """MapPoint descriptors are pinned by the scene builder (the reference chooses between equally good ones by KeyFrame
    address), so two runs of the reference must agree — otherwise the comparison above would be meaningless."""
    a = slam_scenarios.run_all("ref", 7, True)
    b = slam_scenarios.run_all("ref", 7, True)
    assert slam_scenarios.compare(a, b) == []


@pytest.mark.parametrize("w,h,nf,seed,cam", [(1241, 376, 2000, 2, (386.1448, 718.856)), (752, 480, 1200, 1000, (47.9064, 435.2047)),
                                              (640, 480, 1000, 5, (386.1448, 718.856))])
def test_reference_compute_stereo_matches_equals_oracle(w, h, nf, seed, cam):
    left, right = gen_synth.stereo_pair(h, w, seed)
    params = (nf, 1.2, 8, 20, 7)
    ol, orr = oracle.OracleExtractor(*params), oracle.OracleExtractor(*params)
    (kl, dl), (kr, dr) = ol(left), orr(right)
    mbf = cam[0]
    mb = np.float32(np.float32(mbf) / np.float32(cam[1]))
    ou, oz, _, _ = oracle.stereo(ol, orr, kl, dl, kr, dr, mbf, mb)
    ru, rz, matched = slam_ref.SlamLib("ref").compute_stereo(params, left, right, kl, dl, kr, dr, mbf, float(mb))
    assert matched > 100
    assert np.array_equal(ru, ou) and np.array_equal(rz, oz)


def test_reference_compute_stereo_matches_equals_oracle_on_adversarial_pairs():
    """The pairs of tests/test_gpu_stereo.py::test_stereo_edge_cases_in_one_batch through the reference's own Frame.cc:
    no right keypoints, no left keypoints, a 2-px dot lattice shifted by 6 px, 0 / 255 noise shifted by 10 px."""
    h, w = 263, 341
    lat = np.zeros((h, w + 6), np.uint8)
    ii, jj = np.meshgrid(np.arange(0, w + 6, 2) // 2, np.arange(0, h, 2) // 2)
    lat[::2, ::2] = (50 + (ii * 24 + jj) % 192).astype(np.uint8)
    noise = (np.random.default_rng(4).integers(0, 2, (h, w + 10)) * 255).astype(np.uint8)
    tex, flat = gen_synth.texture(h, w, 3), np.full((h, w), 90, np.uint8)
    pairs = [(tex, flat), (flat, tex), (lat[:, :-6], lat[:, 6:]), (noise[:, :-10], noise[:, 10:])]
    params = (500, 1.2, 8, 20, 7)
    mbf = 386.1448
    mb = np.float32(np.float32(mbf) / np.float32(718.856))
    lib = slam_ref.SlamLib("ref")
    for i, (l, r) in enumerate(pairs):
        l, r = np.ascontiguousarray(l), np.ascontiguousarray(r)
        ol, orr = oracle.OracleExtractor(*params), oracle.OracleExtractor(*params)
        (kl, dl), (kr, dr) = ol(l), orr(r)
        ou, oz, _, kept = oracle.stereo(ol, orr, kl, dl, kr, dr, mbf, mb)
        if len(kl) == 0:
            assert len(ou) == 0
            continue
        ru, rz, matched = lib.compute_stereo(params, l, r, kl, dl, kr, dr, mbf, float(mb))
        assert np.array_equal(ru, ou) and np.array_equal(rz, oz), i
        assert matched == kept and (matched == 0) == (i == 0)


def test_descriptor_distance_is_the_reference_popcount():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b[:50] = a[:50]
    b[50:100, :4] = a[50:100, :4]
    L = slam_ref.SlamLib("ref").L
    S = slam_ref.SlamLib("shim_cpu").L
    want = np.unpackbits(a ^ b, axis=1).sum(1)
    for i in range(len(a)):
        r = L.slam_descriptor_distance(slam_ref._ptr(a[i]), slam_ref._ptr(b[i]))
        assert r == want[i] == oracle.hamming(a[i], b[i]) == S.slam_descriptor_distance(slam_ref._ptr(a[i]), slam_ref._ptr(b[i]))


def test_reference_stereo_frame_constructor_grid_equals_oracle_grid():
    """Frame::Frame(stereo) of the reference, whole (two extractor threads, ComputeStereoMatches, ComputeImageBounds,
    AssignFeaturesToGrid): the grid it builds for ITS keypoints equals the oracle's grid restatement."""
    import refactored_orb_slam2_b200 as pkg
    left, right = gen_synth.stereo_pair(240, 416, 11)
    out = slam_ref.SlamLib("ref").stereo_frame((500, 1.2, 8, 20, 7), left, right, 718.856, 718.856, 208.0, 120.0, 386.1448)
    assert len(out["keys"]) > 400 and (out["u_right"] >= 0).sum() > 50
    geo = pkg.grid_geometry(0.0, 416.0, 0.0, 240.0)
    o = oracle.OracleFeatures(out["keys"], out["desc"], out["u_right"], geo)
    start, items = o.grid()
    assert np.array_equal(start, out["cell_start"]) and np.array_equal(items, out["cell_items"])


def test_reference_compute_distinctive_descriptors_equals_oracle():
    """Row f4: the reference's own MapPoint::ComputeDistinctiveDescriptors on MapPoints observed from 12 keyframes; the oracle's
    batched restatement, fed the observation descriptors in the reference's own visiting order, must choose the same one."""
    lib = slam_ref.SlamLib("ref")
    W = slam_ref.World(17, npoints=400, clutter=60, flips=70)
    S = slam_ref.Scene(lib, W)
    views = [W.observe(slam_ref._pose(0.002 * i, -0.01 * i, 0.0, [-0.12 * i, 0.0, -0.1 * i])) for i in range(12)]
    kfs = [S.add_keyframe(S.add_frame(V)) for V in views]
    obs = {p: o for p, o in slam_ref.observations_from_views(views).items() if len(o) >= 3}
    S.add_mappoints(obs)
    rows, offsets, chosen = [], [0], []
    for mp in range(S.n_mps):
        buf, out = np.zeros((16, 32), np.uint8), np.zeros(32, np.uint8)
        n = lib.L.slam_mappoint_distinctive(S.h, mp, slam_ref._ptr(buf), 16, slam_ref._ptr(out))
        rows.append(buf[:n]); offsets.append(offsets[-1] + n); chosen.append(out)
    rows = np.concatenate(rows)
    idx, med = oracle.distinctive(rows, offsets)
    assert S.n_mps > 200 and max(np.diff(offsets)) >= 10
    for g in range(S.n_mps):
        assert np.array_equal(rows[offsets[g] + idx[g]], chosen[g]), g
    S.close()


def test_shim_feature_cache_is_safe_under_concurrent_eviction():
    """Tracking, LocalMapping and LoopClosing call the drop-in ORBmatcher at the same time.  Four host threads search 60
    keyframes with a cache of two device sets, each starting at another keyframe: every search must see its own feature
    set alive for the whole call although the other threads keep evicting it (round 1 returned a raw pointer into the
    cache).  The test backend audits lifetimes: a set destroyed under a running search, or searched after its death, counts."""
    lib = slam_ref.SlamLib("shim_cpu")
    W = slam_ref.World(23, npoints=260, clutter=40)
    S = slam_ref.Scene(lib, W)
    views, kfs, poses = [], [], []
    for i in range(60):
        T = slam_ref._pose(0.001 * i, -0.004 * i, 0.0, [-0.05 * i, 0.0, -0.03 * i])
        V = W.observe(T)
        kfs.append(S.add_keyframe(S.add_frame(V)))
        views.append(V)
        poses.append(np.asarray(T, np.float32))
    S.add_mappoints({p: o for p, o in slam_ref.observations_from_views(views[:6]).items()})
    mps = np.arange(S.n_mps, dtype=np.int32)
    Scws = np.ascontiguousarray(np.stack(poses), np.float32)
    import ctypes
    lib.L.slam_stress_sim3_projection.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.L.slam_stress_sim3_projection.restype = None
    lib.L.orbgpu_oracle_backend_violations.restype = ctypes.c_int
    want = np.zeros((1, 60), np.int32)
    lib.L.slam_stress_sim3_projection(S.h, slam_ref._ptr(Scws), slam_ref._ptr(mps), len(mps), 10, 0.75, 1, 1, slam_ref._ptr(want))
    assert want.sum() > 500
    lib.L.orbgpu_shim_set_feature_cache_capacity(2)          # every search evicts: the hazard no longer needs a long search to show
    got = np.zeros((4, 60), np.int32)
    lib.L.slam_stress_sim3_projection(S.h, slam_ref._ptr(Scws), slam_ref._ptr(mps), len(mps), 10, 0.75, 4, 8, slam_ref._ptr(got))
    S.close()
    assert np.array_equal(got, np.repeat(want, 4, axis=0))
    lib.L.orbgpu_shim_clear_feature_cache()
    lib.L.orbgpu_shim_set_feature_cache_capacity(48)
    assert lib.L.orbgpu_oracle_backend_violations() == 0
