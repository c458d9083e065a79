import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import refactored_orb_slam2_b200 as pkg
names = ["POPC", "LOP3", "VIMNMX3.S16x2", "VIMNMX.S16x2", "VIADD.16x2", "PRMT", "IMAD", "VIMNMX3.S32", "IADD3(2 adds)"]
for m, n in enumerate(names):
    v = ctypes.c_double()
    pkg._lib.check(pkg.lib().orbgpu_measure_pipe(-1, m, ctypes.byref(v)))
    print("%-16s %.3e lane-ops/s  = %.1f lanes/clk/SM @1.965GHz" % (n, v.value, v.value / 148 / 1.965e9))
