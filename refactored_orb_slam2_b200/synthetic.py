"""Deterministic synthetic inputs for the ORB front-end (bench / test inputs; numpy only, no CUDA and no oracle code).

No dataset ships with the reference, so every test and bench input is
generated here from numpy's PCG64 with fixed seeds:

* texture(h, w, seed)        multi-octave Gaussian-filtered white noise, min-max normalised to u8
                             (gives ~10 k FAST candidates on a KITTI-shaped level 0 and pushes
                             most 30x30 cells to the minThFAST fallback).
* stereo_pair(h, w, seed)    left = texture, right = left warped by a piecewise-planar disparity
                             field in [2, 64] px (fractional parts included) + +-2 grey levels.
* frame_batch(n, h, w, seed) n distinct frames cut from one large texture canvas (fast enough to
                             build the 4096-frame sweep of BASELINE.json config 5).
* descriptor_bank / queries_from_bank   BASELINE.json config 4 (2000 x 1M brute-force matching).
"""
import numpy as np

try:  # scipy is in the image; keep a pure-numpy fallback so the generator never disappears
    from scipy.ndimage import gaussian_filter as _gf

    def _gauss(a, s):
        return _gf(a, s, mode="reflect")
except Exception:  # pragma: no cover
    def _gauss(a, s):
        r = int(4 * s + 0.5)
        k = np.exp(-0.5 * (np.arange(-r, r + 1) / s) ** 2)
        k /= k.sum()
        p = np.pad(a, r, mode="symmetric")
        p = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 0, p)
        return np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 1, p)


def texture(h, w, seed):
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), np.float64)
    for sigma, wt in ((1, 0.5), (3, 1.0), (9, 1.5), (25, 2.0)):
        acc += wt * sigma * _gauss(rng.standard_normal((h, w)), sigma)
    acc -= acc.min()
    acc /= max(acc.max(), 1e-12)
    return np.ascontiguousarray((acc * 255.0 + 0.5).astype(np.uint8))


def _warp_rows(img, disp):
    """right(x, y) = left(x + disp(x, y), y) with linear interpolation along x."""
    h, w = img.shape
    xs = np.arange(w, dtype=np.float64)[None, :] + disp
    xs = np.clip(xs, 0, w - 1)
    x0 = np.floor(xs).astype(np.int64)
    x1 = np.minimum(x0 + 1, w - 1)
    f = xs - x0
    rows = np.arange(h)[:, None]
    return img[rows, x0] * (1 - f) + img[rows, x1] * f


def stereo_pair(h, w, seed):
    left = texture(h, w, seed)
    rng = np.random.default_rng(seed + 7919)
    # piecewise-planar disparity: a few vertical strips, each d = a + b*y, clipped to [2, 64]
    nstrips = 5
    edges = np.linspace(0, w, nstrips + 1).astype(int)
    disp = np.zeros((h, w))
    y = np.arange(h, dtype=np.float64)[:, None]
    for s in range(nstrips):
        a = rng.uniform(4, 56)
        b = rng.uniform(-8, 8) / h
        disp[:, edges[s]:edges[s + 1]] = a + b * y
    disp = np.clip(disp, 2.0, 64.0)
    right = _warp_rows(left.astype(np.float64), disp)
    right += rng.integers(-2, 3, size=right.shape)
    right = np.clip(np.rint(right), 0, 255).astype(np.uint8)
    return left, np.ascontiguousarray(right)


def frame_batch(n, h, w, seed, stereo=False, canvas=None):
    """n frames (or n stereo pairs) cropped from one texture canvas at seeded offsets.
    Stereo right frames are the same crop shifted by an integer disparity in [4, 60] plus noise."""
    rng = np.random.default_rng(seed)
    ch, cw = max(2 * h, 1024), max(2 * w, 2048)
    if canvas is None:
        canvas = texture(ch, cw, seed ^ 0x5EED)
    ch, cw = canvas.shape
    left = np.empty((n, h, w), np.uint8)
    right = np.empty((n, h, w), np.uint8) if stereo else None
    for i in range(n):
        oy = int(rng.integers(0, ch - h + 1))
        ox = int(rng.integers(64, cw - w + 1))
        left[i] = canvas[oy:oy + h, ox:ox + w]
        if stereo:
            d = int(rng.integers(4, 61))
            r = canvas[oy:oy + h, ox - d:ox - d + w].astype(np.int16)
            r += rng.integers(-2, 3, size=r.shape, dtype=np.int16)
            right[i] = np.clip(r, 0, 255).astype(np.uint8)
    return (left, right) if stereo else left


def descriptor_bank(n, seed=4):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def _flip_bits(row, k, rng):
    out = row.copy()
    for b in rng.choice(256, size=k, replace=False):
        out[b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def queries_from_bank(bank, nq, seed=5):
    """Queries = bank rows with k ~ U{0..60} flipped bits; for half of them a decoy with
    k + U{0..20} flips is planted in another bank row (exercises ratio 0.7 / TH_LOW 50 both ways).
    Returns (queries, bank_with_decoys, source_rows)."""
    rng = np.random.default_rng(seed)
    bank = bank.copy()
    nb = bank.shape[0]
    src = rng.choice(nb, size=nq, replace=False)
    spare = np.setdiff1d(rng.choice(nb, size=min(nb, 3 * nq), replace=False), src)
    q = np.empty((nq, 32), np.uint8)
    for i, r in enumerate(src):
        k = int(rng.integers(0, 61))
        q[i] = _flip_bits(bank[r], k, rng)
        if i % 2 == 0 and i // 2 < len(spare):
            k2 = min(256, k + int(rng.integers(0, 21)))
            bank[spare[i // 2]] = _flip_bits(q[i], k2, rng)
    return q, bank, src
