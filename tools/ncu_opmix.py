"""Executed-instruction mix by opcode and by barrier-delimited phase from an ncu source-page CSV."""
import collections
import csv
import re
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
tot = sum(int(r[ie]) for r in data)
op = collections.Counter(); samp = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ia])
    o = m.group(2) if m else r[ia].split()[0]
    op[o] += int(r[ie]); samp[o] += int(r[isamp])
print('total warp instr', tot)
for o, c in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print('%-10s %12d %5.1f%%  samples %6d' % (o, c, 100 * c / tot, samp[o]))
cum = 0
for i, r in enumerate(data):
    cum += int(r[ie])
    if 'BAR' in r[ia] or 'EXIT' in r[ia] or 'SYNCS.PHASECHK' in r[ia]:
        print(i, r[ia].strip()[:60], 'cum %.1f%%' % (100 * cum / tot))
