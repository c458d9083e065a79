"""Frame::UndistortKeyPoints / ComputeImageBounds (Frame.cc:419-475): the oracle's restatement of cv::undistortPoints against
cv2 itself (CPU), and the device kernel against both (GPU), bit for bit, on the reference's own calibration files."""
import numpy as np
import pytest

import oracle

cv2 = pytest.importorskip("cv2")

# (fx, fy, cx, cy, [k1 k2 p1 p2 (k3)], width, height) from Examples/Monocular/TUM1.yaml, TUM2.yaml, EuRoC.yaml and TUM3 (no distortion)
CAMERAS = [
    (517.306408, 516.469215, 318.643040, 255.313989, [0.262383, -0.953104, -0.005358, 0.002628, 1.163314], 640, 480),
    (520.908620, 521.007327, 325.141442, 249.701764, [0.231222, -0.784899, -0.003257, -0.000105, 0.917205], 640, 480),
    (458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], 752, 480),
    (535.4, 539.2, 320.1, 247.6, [0.0, 0.0, 0.0, 0.0], 640, 480),
]


def points(w, h, n, seed):
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], axis=1).astype(np.float32)
    p[:4] = [[0, 0], [w, 0], [0, h], [w, h]]                      # the corners of ComputeImageBounds
    p[4:8] = np.rint(p[4:8])
    return p


def cv2_undistort(p, cam):
    fx, fy, cx, cy, dist = cam[:5]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    return cv2.undistortPoints(p.reshape(-1, 1, 2), K, np.array(dist, np.float32).reshape(-1, 1), None, K).reshape(-1, 2)


@pytest.mark.parametrize("cam", CAMERAS, ids=["TUM1", "TUM2", "EuRoC", "TUM3_nodist"])
def test_oracle_equals_cv2(cam):
    p = points(cam[5], cam[6], 100000, 3)
    got = oracle.undistort_points(p, *[np.float32(v) for v in cam[:4]], cam[4])
    want = cv2_undistort(p, cam)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("cam", CAMERAS, ids=["TUM1", "TUM2", "EuRoC", "TUM3_nodist"])
def test_device_equals_oracle_and_cv2(built, cam):
    import refactored_orb_slam2_b200 as pkg
    p = points(cam[5], cam[6], 200000, 4)
    camera = pkg.Camera.make(*cam[:4], cam[4])
    got = pkg.undistort_points(p, camera)
    assert np.array_equal(got.view(np.uint32), oracle.undistort_points(p, *[np.float32(v) for v in cam[:4]], cam[4]).view(np.uint32))
    assert np.array_equal(got.view(np.uint32), cv2_undistort(p, cam).view(np.uint32))
    # keypoint records: everything but pt is carried over
    k = np.zeros(5000, pkg.KP_DTYPE)
    k["x"], k["y"] = p[:5000, 0], p[:5000, 1]
    k["size"], k["angle"], k["response"], k["octave"], k["class_id"] = 31, np.arange(5000) % 360, 20, np.arange(5000) % 8, -1
    ku = pkg.UndistortKeyPoints(k, camera)
    assert np.array_equal(np.stack([ku["x"], ku["y"]], axis=1).view(np.uint32), got[:5000].view(np.uint32))
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(ku[f], k[f])
    # ComputeImageBounds
    b = pkg.ComputeImageBounds(cam[5], cam[6], camera)
    c = cv2_undistort(np.array([[0, 0], [cam[5], 0], [0, cam[6]], [cam[5], cam[6]]], np.float32), cam)
    assert b == (min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1]))


@pytest.mark.gpu
def test_feature_set_from_extractor_undistorts_on_the_device(built):
    """Distorted TUM1 input: extraction -> device-side UndistortKeyPoints -> grid, no host round trip; equals the host path
    (download, cv2.undistortPoints, upload) in keypoints and in the grid AssignFeaturesToGrid builds from them."""
    import gen_synth
    import refactored_orb_slam2_b200 as pkg
    cam = CAMERAS[0]
    img = gen_synth.texture(480, 640, 31)
    ex = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=1)
    kps, desc, counts = ex.extract_batch(img[None])
    n = counts[0]
    camera = pkg.Camera.make(*cam[:4], cam[4])
    minx, maxx, miny, maxy = pkg.ComputeImageBounds(640, 480, camera)
    geo = pkg.grid_geometry(minx, maxx, miny, maxy)
    dev = pkg.FrameFeatures.from_extractor(ex, 0, geo, camera=camera)
    k_dev, _, d_dev = dev.download()
    un = cv2_undistort(np.stack([kps[0, :n]["x"], kps[0, :n]["y"]], axis=1), cam)
    assert np.array_equal(np.stack([k_dev["x"], k_dev["y"]], axis=1).view(np.uint32), un.view(np.uint32))
    assert np.array_equal(d_dev, desc[0, :n]) and np.array_equal(k_dev["octave"], kps[0, :n]["octave"])
    host_keys = kps[0, :n].copy()
    host_keys["x"], host_keys["y"] = un[:, 0], un[:, 1]
    host = pkg.FrameFeatures(host_keys, desc[0, :n], None, geo)
    s1, i1 = dev.grid()
    s2, i2 = host.grid()
    assert np.array_equal(s1, s2) and np.array_equal(i1, i2)
    assert (un != np.stack([kps[0, :n]["x"], kps[0, :n]["y"]], axis=1)).any()
