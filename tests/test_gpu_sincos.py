"""computeOrbDescriptor's cos / sin (ORBextractor.cc:102-106): the descriptor kernel's device sincos against this host's libm
cosf / sinf on EVERY float32 of [0, 2 pi] (and a margin above it: angle = fastAtan2 degrees * (pi / 180) as float can exceed
2 pi by a few ulp).  ~1.09 G arguments, compared through per-chunk 64-bit checksums computed on both sides."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

LAST = 0x40C91000                      # 2 pi = 0x40C90FDB; the sweep covers [+0, 0x40C91000]
CHUNK_LOG2 = 20


def test_device_sincos_equals_libm_on_every_float_of_the_angle_range(built):
    import refactored_orb_slam2_b200 as pkg
    count = LAST + 1
    chunks = (count + (1 << CHUNK_LOG2) - 1) >> CHUNK_LOG2
    dev = np.zeros(chunks, np.uint64)
    pkg._lib.check(pkg.lib().orbgpu_selftest_sincos(0, count, CHUNK_LOG2, pkg._lib.ptr(dev), None, None, -1))
    host = np.zeros(chunks, np.uint64)
    L = oracle.lib()
    L.orb_oracle_sincos_checksum.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
    L.orb_oracle_sincos_checksum.restype = None
    L.orb_oracle_sincos_checksum(0, count, CHUNK_LOG2, host.ctypes.data_as(C.c_void_p), os.cpu_count() or 1)
    bad = np.nonzero(dev != host)[0]
    if len(bad):                        # name the first differing argument
        first = int(bad[0]) << CHUNK_LOG2
        n = min(1 << CHUNK_LOG2, count - first)
        s, c = np.zeros(n, np.float32), np.zeros(n, np.float32)
        pkg._lib.check(pkg.lib().orbgpu_selftest_sincos(first, n, CHUNK_LOG2, None, pkg._lib.ptr(s), pkg._lib.ptr(c), -1))
        x = (np.arange(n, dtype=np.uint32) + np.uint32(first)).view(np.float32)
        hs, hc = np.sin(x.astype(np.float32)), np.cos(x.astype(np.float32))
        k = int(np.nonzero((s.view(np.uint32) != hs.view(np.uint32)) | (c.view(np.uint32) != hc.view(np.uint32)))[0][0]) if n else 0
        pytest.fail("%d of %d chunks differ; first argument bits 0x%08x: device (%r, %r)" % (len(bad), chunks, first + k, s[k], c[k]))
    assert host.any()


def test_value_dump_matches_libm_on_the_fastatan2_grid(built):
    """The raw-value path of the hook on the angles IC_Angle can actually produce near the quadrant boundaries."""
    import refactored_orb_slam2_b200 as pkg
    for first in (0x3FC90FDB - 4096, 0x40490FDB - 4096, 0x4096CBE4 - 4096, 0x40C90FDB - 4096, 0):      # around pi/2, pi, 3 pi/2, 2 pi, 0
        n = 8192
        s, c = np.zeros(n, np.float32), np.zeros(n, np.float32)
        pkg._lib.check(pkg.lib().orbgpu_selftest_sincos(first, n, 20, None, pkg._lib.ptr(s), pkg._lib.ptr(c), -1))
        for i in range(0, n, 97):
            bits = np.uint32(first + i)
            hc, hs = oracle.sincos(float(bits.view(np.float32)))
            assert np.float32(hs).view(np.uint32) == s[i].view(np.uint32) and np.float32(hc).view(np.uint32) == c[i].view(np.uint32), hex(first + i)
