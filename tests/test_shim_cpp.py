"""The C++ drop-in shim (shim/ORBextractor.{h,cc}, shim/FrameStereo.h): it must build against the
OpenCV-compat stub without a GPU, keep the reference's public interface, and — on the GPU box —
produce the same bytes as the ctypes path (hence as the oracle)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim")


def test_shim_builds_and_keeps_the_reference_interface(built):
    subprocess.check_call(["make", "-s", "-C", SHIM])
    assert os.path.exists(os.path.join(SHIM, "shim_demo"))
    hdr = open(os.path.join(SHIM, "ORBextractor.h")).read()
    # the public surface of ORBextractor.h:47-75 in the reference
    for needle in ["class ORBextractor", "enum { HARRIS_SCORE = 0, FAST_SCORE = 1 }",
                   "ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)",
                   "void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints",
                   "GetLevels()", "GetScaleFactor()", "GetScaleFactors()", "GetInverseScaleFactors()",
                   "GetScaleSigmaSquares()", "GetInverseScaleSigmaSquares()", "std::vector<cv::Mat> mvImagePyramid"]:
        assert needle in hdr, needle
    assert re.search(r"namespace ORB_SLAM2", hdr)


@pytest.mark.gpu
def test_shim_matches_python_path(built, tmp_path):
    import gen_synth
    import refactored_orb_slam2_b200 as pkg
    subprocess.check_call(["make", "-s", "-C", SHIM])
    w, h, nf = 752, 480, 1200
    left, right = gen_synth.stereo_pair(h, w, 77)
    left.tofile(tmp_path / "l.raw")
    right.tofile(tmp_path / "r.raw")
    out = tmp_path / "out.bin"
    subprocess.check_call([os.path.join(SHIM, "shim_demo"), str(tmp_path / "l.raw"), str(tmp_path / "r.raw"),
                           str(w), str(h), str(nf), str(out)])
    raw = out.read_bytes()
    n = int(np.frombuffer(raw, np.int32, 1)[0])
    off = 4
    kps = np.frombuffer(raw, pkg.KP_DTYPE, n, off); off += 28 * n
    desc = np.frombuffer(raw, np.uint8, 32 * n, off).reshape(n, 32); off += 32 * n
    u = np.frombuffer(raw, np.float32, n, off); off += 4 * n
    z = np.frombuffer(raw, np.float32, n, off); off += 4 * n
    levels, tw, th = np.frombuffer(raw, np.int32, 3, off); off += 12
    top = np.frombuffer(raw, np.uint8, tw * th, off).reshape(th, tw)
    exl, exr = pkg.ORBextractor(nf, 1.2, 8, 20, 7), pkg.ORBextractor(nf, 1.2, 8, 20, 7)
    f = pkg.Frame(left, right, exl, exr, np.float32(386.1448), np.float32(718.856))
    assert n == f.N
    for name in kps.dtype.names:
        assert np.array_equal(kps[name], f.mvKeys[name]), name
    assert np.array_equal(desc, f.mDescriptors)
    assert np.array_equal(u, f.mvuRight) and np.array_equal(z, f.mvDepth)
    assert levels == 8 and np.array_equal(top, exl.pyramid_level(7))
