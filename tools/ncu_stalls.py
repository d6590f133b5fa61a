"""Developer tool: stall-reason totals and the hottest instructions of one kernel from `ncu -i X.ncu-rep --page source --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
body = rows[2:]
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = totinst = 0
agg = {s: 0 for s in stalls}
recs = []
for r in body:
    try:
        n = int(r[ix['Instructions Executed']])
        sm = int(r[ix['# Samples']])
    except (ValueError, IndexError):
        continue
    tot += sm
    totinst += n
    for s in stalls:
        try:
            agg[s] += int(r[ix[s]])
        except ValueError:
            pass
    recs.append((sm, n, r[ix['Source']].strip()[:90], {s: int(r[ix[s]]) for s in stalls if r[ix[s]] not in ('', '0')}))
print('samples', tot, 'warp instructions', totinst)
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f'{s:24s} {100 * v / tot:5.1f}%')
print('--- top instructions by samples')
for sm, n, src, st in sorted(recs, key=lambda t: -t[0])[:top]:
    print(f'{100 * sm / tot:5.2f}% n={n:8d} {src}  {dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])}')
