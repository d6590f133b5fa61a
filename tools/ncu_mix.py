"""Instruction mix and stall-sample summary of one kernel from `ncu --page source --csv` output (file given as argv[1])."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
src, ex, samp = ix['Source'], ix['Instructions Executed'], ix['# Samples']
c, cs = Counter(), Counter()
tot = tots = 0
body = rows[2:]
for r in body:
    try:
        n, sm = int(r[ex]), int(r[samp])
    except (ValueError, IndexError):
        continue
    t = r[src].split()
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    c[op] += n; cs[op] += sm; tot += n; tots += sm
print('warp instructions', tot, 'samples', tots)
for k, v in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f'{k:12s} {v:10d} {100*v/tot:5.1f}%  samples {100*cs[k]/max(tots,1):5.1f}%')
if len(sys.argv) > 3:   # segments between barriers
    seg = 0; segn = Counter(); segs = Counter(); names = {}
    for r in body:
        try: n, sm = int(r[ex]), int(r[samp])
        except (ValueError, IndexError): continue
        segn[seg] += n; segs[seg] += sm
        t = r[src]
        if 'BAR.SYNC' in t or 'SYNCS' in t or 'UTCBAR' in t or 'UTCHMMA' in t and False:
            names[seg] = t.strip()[:60]; seg += 1
    for k in sorted(segn): print(k, segn[k], f'{100*segs[k]/max(tots,1):5.1f}%', names.get(k, ''))
