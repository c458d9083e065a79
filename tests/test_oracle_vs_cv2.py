"""Pins oracle/oracle_cpu.cpp: every OpenCV primitive against the real cv2, the whole extractor
against the independent cv2-driven restatement (oracle/oracle_cv2.py), and the in-source constants
of the reference (quota / umax formulas, pattern table checksum).  CPU only."""
import hashlib

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import gen_synth
import oracle
from oracle_cv2 import Cv2Extractor, load_pattern

cv2.setNumThreads(1)
SHAPES = [(640, 480), (1241, 376), (752, 480), (179, 134), (97, 61)]


def test_pattern_table_checksum():
    # sha256 of the 1024 int8 pattern values taken from ORBextractor.cc:148-405 at survey time
    p = load_pattern().astype(np.int8)
    assert hashlib.sha256(p.tobytes()).hexdigest() == "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"
    assert np.abs(p).max() == 13


def test_constructor_tables():
    e = oracle.OracleExtractor(1000, 1.2, 8, 20, 7)
    assert e.quota.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert e.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert oracle.OracleExtractor(2000).quota.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle.OracleExtractor(1200).quota.tolist() == [261, 217, 181, 151, 126, 105, 87, 72]
    assert np.allclose(e.scale, 1.2 ** np.arange(8), rtol=1e-6)
    c = Cv2Extractor(1000)
    assert c.quota == e.quota.tolist() and c.umax == e.umax.tolist()
    assert [float(s) for s in c.scale] == e.scale.tolist()


@pytest.mark.parametrize("w,h", SHAPES + [(3840, 2160)])
def test_resize_bit_exact(w, h):
    rng = np.random.default_rng(w)
    cur = rng.integers(0, 256, (h, w), dtype=np.uint8) if w > 2000 else gen_synth.texture(h, w, 1)
    for l in range(1, 8):
        s = np.float32(1) / np.float32(1.2) ** l
        dw, dh = int(np.rint(np.float32(w) * np.float32(s))), int(np.rint(np.float32(h) * np.float32(s)))
        if dw < 8 or dh < 8:
            break
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, oracle.resize(cur, dw, dh)), "level %d" % l
        cur = ref


@pytest.mark.parametrize("w,h", SHAPES + [(50, 40), (7, 9)])
def test_gaussian_blur_bit_exact(w, h):
    rng = np.random.default_rng(h)
    for img in (rng.integers(0, 256, (h, w), dtype=np.uint8), gen_synth.texture(max(h, 32), max(w, 32), 3)[:h, :w].copy()):
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, oracle.blur(img))


def test_fast_atan2_bit_exact():
    rng = np.random.default_rng(0)
    ys, xs = rng.integers(-3000000, 3000000, 20000), rng.integers(-3000000, 3000000, 20000)
    pts = list(zip(ys, xs)) + [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (5, 5), (-5, 5), (5, -5), (-5, -5)]
    for y, x in pts:
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(oracle.fast_atan2(y, x)), (y, x)


@pytest.mark.parametrize("th", [7, 12, 20])
def test_fast_nms_bit_exact_including_order(th):
    rng = np.random.default_rng(th)
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for (w, h) in [(37, 38), (31, 32), (320, 240), (7, 7), (8, 6), (100, 9), (6, 40)]:
        for img in (rng.integers(0, 256, (h, w), dtype=np.uint8), gen_synth.texture(max(h, 32), max(w, 32), 5)[:h, :w].copy()):
            ref = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in det.detect(img)], np.int32).reshape(-1, 3)
            assert np.array_equal(ref, oracle.fast(img, th)), (w, h)


def test_border_is_reflect101():
    img = gen_synth.texture(100, 120, 2)
    e = oracle.OracleExtractor(300)
    e(img)
    assert np.array_equal(e.level(0, 1), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("w,h,nf,seed", [(640, 480, 1000, 1), (416, 240, 500, 11), (333, 251, 300, 5)])
def test_full_extractor_matches_cv2_restatement(w, h, nf, seed):
    img = gen_synth.texture(h, w, seed)
    e, c = oracle.OracleExtractor(nf), Cv2Extractor(nf)
    k, d = e(img)
    k2, d2 = c(img)
    for l in range(8):
        assert np.array_equal(e.level(l, 1), c.pyr[l]), "pyramid level %d" % l
        assert np.array_equal(e.stage(l, 0), np.array(c.trace[l][0], np.int32).reshape(-1, 3)), "candidates level %d" % l
        assert np.array_equal(e.stage(l, 1), np.array(c.trace[l][1], np.int32).reshape(-1, 3)), "selection level %d" % l
    assert len(k) == len(k2) and len(k) >= nf
    for f in k.dtype.names:
        assert np.array_equal(k[f], k2[f]), f
    assert np.array_equal(d, d2)


def test_sincos_polynomial_matches_libm():
    """The double-precision sincosf algorithm the GPU uses (describe.cu) vs this host's libm,
    on a strided sweep of all float32 angles in [0, 2*pi]."""
    hi = np.float32(6.2831855).view(np.uint32)
    bits = np.arange(0, int(hi) + 1, 9973, dtype=np.uint32)
    ang = bits.view(np.float32)
    x = ang.astype(np.float64)
    top = (bits >> 20) & 0x7ff
    small = top < ((np.float32(0.7853982).view(np.uint32) >> 20) & 0x7ff)
    n = ((x * float.fromhex("0x1.45F306DC9C883p+23")).astype(np.int32).astype(np.int64) + 0x800000) >> 24
    n = np.where(small, 0, n)
    xr = np.where(small, x, x - n * float.fromhex("0x1.921FB54442D18p0"))
    sgn = np.where(((n & 3) == 1) | ((n & 3) == 2), -1.0, 1.0)
    neg = (n & 2) != 0
    xs, x2 = xr * sgn, xr * xr
    C = [1.0, float.fromhex("-0x1.ffffffd0c621cp-2"), float.fromhex("0x1.55553e1068f19p-5"),
         float.fromhex("-0x1.6c087e89a359dp-10"), float.fromhex("0x1.99343027bf8c3p-16")]
    S = [float.fromhex("-0x1.555545995a603p-3"), float.fromhex("0x1.1107605230bc4p-7"), float.fromhex("-0x1.994eb3774cf24p-13")]

    def poly(cosine):
        x3 = xs * x2
        s = (xs + x3 * S[0]) + (x3 * x2) * (S[1] + x2 * S[2])
        x4 = x2 * x2
        sg = np.where(neg, -1.0, 1.0)
        c = ((sg * C[0] + x2 * (sg * C[1])) + x4 * (sg * C[2])) + (x4 * x2) * (sg * C[3] + x2 * (sg * C[4]))
        return np.where(cosine, c, s).astype(np.float32)

    sin_p, cos_p = poly((n & 1) != 0), poly((n & 1) == 0)
    tiny = top < ((np.float32(2.0 ** -12).view(np.uint32) >> 20) & 0x7ff)
    sin_p = np.where(tiny, ang, sin_p)
    cos_p = np.where(tiny, np.float32(1), cos_p)
    ref = np.array([oracle.sincos(a) for a in ang], np.float32)
    assert np.array_equal(cos_p, ref[:, 0]) and np.array_equal(sin_p, ref[:, 1])
