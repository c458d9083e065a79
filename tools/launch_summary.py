"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum ... --csv --log-file X)."""
import collections
import csv
import sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0].replace("unnamed>::", "").replace("void ", "").replace("orbgpu::<", "").replace("orbgpu::", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
print("%-44s %5s %12s %10s %7s" % ("kernel", "n", "total us", "avg us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-44s %5d %12.1f %10.1f %6.1f%%" % (k[:44], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
