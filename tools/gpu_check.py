"""Stage-by-stage GPU-vs-oracle report (diagnostic; prints every mismatch instead of stopping)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import gen_synth
import oracle
import refactored_orb_slam2_b200 as pkg


def cmp(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        print("   %-12s SHAPE %s vs %s" % (name, a.shape, b.shape))
        return False
    bad = int((a != b).sum())
    print("   %-12s %s (%d / %d differ)" % (name, "ok" if bad == 0 else "MISMATCH", bad, a.size))
    return bad == 0


def check_extract(w, h, nf, seed, kind="texture"):
    print("== extract %dx%d nf=%d seed=%d %s" % (w, h, nf, seed, kind))
    img = gen_synth.texture(h, w, seed) if kind == "texture" else np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    params = (nf, 1.2, 8, 20, 7)
    o = oracle.OracleExtractor(*params)
    ok, od = o(img)
    g = pkg.ORBextractor(*params, max_width=w, max_height=h, max_batch=1)
    t = time.time()
    try:
        gk, gd = g(img)
    except Exception as e:
        print("   GPU extract FAILED:", e)
        gk = gd = None
    print("   gpu call %.1f ms" % ((time.time() - t) * 1e3))
    allok = True
    for l in range(8):
        print("  level %d %s" % (l, o.level_size(l)))
        allok &= cmp("pyramid", g.pyramid_level(l, with_border=True), o.level(l, 1))
        try:
            c0 = g.stage(l, 0)
            allok &= cmp("candidates", c0, o.stage(l, 0))
            allok &= cmp("selected", g.stage(l, 1), o.stage(l, 1))
        except Exception as e:
            print("   stage FAILED:", e); allok = False
        ob = o.level(l, 2)
        if ob is not None:
            allok &= cmp("blurred", g.blurred_level(l), ob)
    if gk is not None:
        print("  keypoints gpu %d oracle %d" % (len(gk), len(ok)))
        if len(gk) == len(ok):
            for f in ok.dtype.names:
                allok &= cmp(f, gk[f], ok[f])
            allok &= cmp("descriptors", gd, od)
            bad = np.nonzero((gd != od).any(axis=1))[0]
            if len(bad):
                print("   first bad rows", bad[:10], "octaves", ok["octave"][bad[:10]], "angles", ok["angle"][bad[:10]], gk["angle"][bad[:10]])
        else:
            allok = False
    print("  RESULT", "PASS" if allok else "FAIL")
    return allok


def check_stereo(w, h, nf, seed):
    print("== stereo %dx%d nf=%d" % (w, h, nf))
    left, right = gen_synth.stereo_pair(h, w, seed)
    params = (nf, 1.2, 8, 20, 7)
    ol, orr = oracle.OracleExtractor(*params), oracle.OracleExtractor(*params)
    (kl, dl), (kr, dr) = ol(left), orr(right)
    mbf, fx = 386.1448, 718.856
    ou, oz, obd, kept = oracle.stereo(ol, orr, kl, dl, kr, dr, mbf, np.float32(mbf / fx))
    g = pkg.ORBextractor(*params, max_width=w, max_height=h, max_batch=2)
    kps, desc, counts = g.extract_batch(np.stack([left, right]))
    ok = cmp("countsL", counts[0], len(kl)) & cmp("countsR", counts[1], len(kr))
    if ok:
        ok &= cmp("descL", desc[0, :len(kl)], dl) & cmp("descR", desc[1, :len(kr)], dr)
    u, z = pkg.ComputeStereoMatches(g, g, mbf, mbf / fx, 0, 1, 1)
    n = len(kl)
    ok &= cmp("uRight", u[0, :n], ou)
    ok &= cmp("depth", z[0, :n], oz)
    print("   oracle matches kept %d / %d ; gpu %d" % (kept, n, int((u[0, :n] >= 0).sum())))
    print("  RESULT", "PASS" if ok else "FAIL")
    return ok


def check_hamming():
    print("== hamming")
    ok = True
    for nb, nq in ((5000, 37), (100000, 2000), (777, 600), (0, 5), (130, 1)):
        bank = gen_synth.descriptor_bank(max(nb, 1), 4)[:nb]
        if nb >= nq and nb > 0:
            q, bank, _ = gen_synth.queries_from_bank(bank, nq, 5)
        else:
            q = gen_synth.descriptor_bank(nq, 9)
        t = time.time()
        idx1, d1, d2 = pkg.DescriptorBank(bank, index_base=1000).top2(q)
        dt = time.time() - t
        oi, o1, o2 = oracle.top2_dense(q, bank, 8) if nb else (np.full(nq, -1), np.full(nq, 256), np.full(nq, 256))
        oi = np.where(oi >= 0, oi + 1000, -1)
        print("  bank %d x %d  (%.1f ms)" % (nb, nq, dt * 1e3))
        ok &= cmp("idx1", idx1, oi) & cmp("d1", d1, o1) & cmp("d2", d2, o2)
    # duplicates / ties
    bank = np.repeat(gen_synth.descriptor_bank(50, 1), 40, axis=0)
    q = bank[::7][:100].copy()
    idx1, d1, d2 = pkg.DescriptorBank(bank).top2(q)
    oi, o1, o2 = oracle.top2_dense(q, bank, 1)
    ok &= cmp("tie idx1", idx1, oi) & cmp("tie d1", d1, o1) & cmp("tie d2", d2, o2)
    # csr
    rng = np.random.default_rng(3)
    train = gen_synth.descriptor_bank(3000, 2)
    q = gen_synth.descriptor_bank(500, 3)
    lens = rng.integers(0, 60, 500)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.integers(0, 3000, off[-1]).astype(np.int32)
    m = pkg.ORBmatcher(0.7)
    match, d1, d2 = m.SearchCandidates(q, train, off, cand)
    oi, o1, o2 = oracle.top2_csr(q, train, off, cand)
    ok &= cmp("csr d1", d1, o1) & cmp("csr d2", d2, o2)
    acc = oracle.accept(o1, o2, 0.7, 50)
    ok &= cmp("csr match", match, np.where(acc, oi, -1))
    a, b = gen_synth.descriptor_bank(1000, 5), gen_synth.descriptor_bank(1000, 6)
    dd = pkg.ORBmatcher.DescriptorDistance(a, b)
    ok &= cmp("distance", dd, np.array([oracle.hamming(x, y) for x, y in zip(a, b)]))
    print("  RESULT", "PASS" if ok else "FAIL")
    return ok


if __name__ == "__main__":
    res = []
    res.append(check_hamming())
    res.append(check_extract(640, 480, 1000, 1))
    res.append(check_extract(1241, 376, 2000, 2))
    res.append(check_extract(333, 251, 300, 5, "noise"))
    res.append(check_stereo(1241, 376, 2000, 2))
    import ctypes
    p, l = ctypes.c_double(), ctypes.c_double()
    pkg.lib().orbgpu_measure_popc_peak(-1, ctypes.byref(p), ctypes.byref(l))
    print("POPC peak %.3e /s  LOP3 peak %.3e /s" % (p.value, l.value))
    print("ALL", res)
    sys.exit(0 if all(res) else 1)
