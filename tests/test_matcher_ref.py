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
