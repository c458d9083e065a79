"""cvtColor(*2GRAY) and remap(INTER_LINEAR) — the oracle's restatements against real cv2 (CPU), and the CUDA
kernels against both (GPU).  Bit-exact."""
import numpy as np
import pytest

import oracle

cv2 = pytest.importorskip("cv2")


def euroc_like_maps(w, h, seed):
    """Smooth undistort-rectify style maps with sub-pixel structure, plus regions that leave the source image."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    k1 = np.float32(-2.8e-7)
    r2 = (xx - w / 2) ** 2 + (yy - h / 2) ** 2
    mx = (xx + (xx - w / 2) * k1 * r2 + np.float32(3.3)).astype(np.float32)
    my = (yy + (yy - h / 2) * k1 * r2 - np.float32(2.6)).astype(np.float32)
    mx += rng.uniform(-0.5, 0.5, mx.shape).astype(np.float32)
    my += rng.uniform(-0.5, 0.5, my.shape).astype(np.float32)
    mx[:4] -= 40; my[:, :4] -= 30; mx[-3:] += 60; my[:, -2:] += 50     # taps outside: BORDER_CONSTANT
    mx[10, 10:20] = np.arange(10, 20) + 0.5 / 32                       # exact .5 ties of cvRound(v * 32)
    return mx, my


@pytest.mark.parametrize("channels,code,cvcode", [(3, 0, "COLOR_BGR2GRAY"), (3, 1, "COLOR_RGB2GRAY"), (4, 0, "COLOR_BGRA2GRAY"), (4, 1, "COLOR_RGBA2GRAY")])
def test_oracle_gray_equals_cv2(channels, code, cvcode):
    img = np.random.default_rng(channels + code).integers(0, 256, (123, 157, channels), dtype=np.uint8)
    img[:2] = 255; img[2:4] = 0
    assert np.array_equal(oracle.cvt_gray(img, code), cv2.cvtColor(img, getattr(cv2, cvcode)))


@pytest.mark.parametrize("w,h", [(752, 480), (333, 251)])
def test_oracle_remap_equals_cv2(w, h):
    src = np.random.default_rng(1).integers(0, 256, (h, w), dtype=np.uint8)
    mx, my = euroc_like_maps(w, h, 2)
    assert np.array_equal(oracle.remap(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR))
    small = np.ascontiguousarray(mx[:100, :200]), np.ascontiguousarray(my[:100, :200])      # output smaller than the source
    assert np.array_equal(oracle.remap(src, *small), cv2.remap(src, *small, cv2.INTER_LINEAR))


@pytest.mark.gpu
@pytest.mark.parametrize("channels,code,cvcode", [(3, 0, "COLOR_BGR2GRAY"), (3, 1, "COLOR_RGB2GRAY"), (4, 0, "COLOR_BGRA2GRAY"), (4, 1, "COLOR_RGBA2GRAY")])
def test_gpu_gray(channels, code, cvcode):
    import refactored_orb_slam2_b200 as pkg
    imgs = np.random.default_rng(7).integers(0, 256, (5, 376, 1241, channels), dtype=np.uint8)      # odd width: tail pixels, unaligned rows
    got = pkg.staging.cvtColorGray(imgs, code)
    for i in range(len(imgs)):
        assert np.array_equal(got[i], cv2.cvtColor(imgs[i], getattr(cv2, cvcode)))
        assert np.array_equal(got[i], oracle.cvt_gray(imgs[i], code))
    one = pkg.staging.cvtColorGray(imgs[0, :7, :3], code)
    assert np.array_equal(one, oracle.cvt_gray(imgs[0, :7, :3], code))


@pytest.mark.gpu
def test_gpu_remap_and_staged_extraction():
    """remap on the device, then the staged frames go through the extractor: same keypoints as extracting cv2's rectified frames."""
    import refactored_orb_slam2_b200 as pkg
    import gen_synth
    w, h = 752, 480
    left, right = gen_synth.stereo_pair(h, w, 21)
    mx, my = euroc_like_maps(w, h, 3)
    got = pkg.staging.remap(np.stack([left, right]), mx, my)
    for i, im in enumerate((left, right)):
        want = cv2.remap(im, mx, my, cv2.INTER_LINEAR)
        assert np.array_equal(got[i], want)
        assert np.array_equal(got[i], oracle.remap(im, mx, my))
    params = (1200, 1.2, 8, 20, 7)
    ex = pkg.ORBextractor(*params, max_width=w, max_height=h, max_batch=2)
    kps, desc, counts = ex.extract_batch(got)
    k, d = oracle.OracleExtractor(*params)(cv2.remap(left, mx, my, cv2.INTER_LINEAR))
    assert counts[0] == len(k) and np.array_equal(kps[0, :len(k)]["x"], k["x"]) and np.array_equal(desc[0, :len(k)], d)
