"""SASS opcode histogram per kernel of liborbgpu.so (cuobjdump -sass): the instruction-level evidence for the Blackwell-native
claims (UTMALDG = TMA tensor copies, UBLKCP = bulk copies, SYNCS = mbarrier, VIMNMX3.S16x2 = packed 3-input min/max,
IDP.4A / IDP.2A = byte dot products, POPC; no HMMA / UTCHMMA: no tensor cores on this path, as north_star requires)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "refactored_orb_slam2_b200/liborbgpu.so"
out = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
        name = re.sub(r"\(anonymous namespace\)::|orbgpu::", "", name).split("(")[0]
        kern = hist.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern is not None:
        kern[m.group(1)] += 1
print("arch: %s" % ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", out)))))
marks = ("UTMALDG", "UBLKCP", "SYNCS", "VIMNMX3", "VIADDMNMX", "IDP", "POPC", "PRMT", "HMMA", "UTCHMMA", "UTCQMMA", "CCTL", "REDUX", "ATOMS")
total = collections.Counter()
for name, h in hist.items():
    n = sum(h.values())
    if name.startswith("k_pipe_probe") or name.startswith("void k_pipe_probe"):
        continue
    total.update(h)
    top = ", ".join("%s %d" % kv for kv in h.most_common(8))
    flagged = ", ".join("%s %d" % (k, v) for k, v in sorted(h.items()) if any(k.startswith(m) for m in marks))
    print("\n%s  (%d instructions)\n  top: %s\n  marked: %s" % (name, n, top, flagged or "-"))
print("\nlibrary totals (probe kernels excluded): " + ", ".join("%s %d" % (k, sum(v for kk, v in total.items() if kk.startswith(k))) for k in marks))
