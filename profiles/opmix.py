#!/usr/bin/env python
"""Dynamic SASS opcode mix per pixel from `ncu --page source --csv` (argv[1]); argv[2] = pixels per launch."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iSamp = hdr.index('# Samples')
tot = 0; by = collections.Counter(); samp = collections.Counter()
for r in rows[2:]:
    try: n = int(r[iE])
    except Exception: continue
    src = r[iS].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', src)
    op = m.group(2) if m else src
    by[op] += n; tot += n; samp[op] += int(r[iSamp])
px = float(sys.argv[2]) / 32
print('total warp-inst', tot, ' per pixel-lane', round(tot / px, 1))
for op, n in by.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print(f'{op:12s} {n/px:8.2f} /px   stall samples {samp[op]}')
