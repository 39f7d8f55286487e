"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scratch/agg_launches.py file.csv [skip_fraction]  (skip the first fraction of launches: warm-up)"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = None
recs = []
for r in rows:
    if len(r) > 5 and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        u = d['Metric Unit']
        v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}.get(u, 1.0)
        recs.append((re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('<unnamed>::', ''), v))
recs = recs[int(len(recs) * skip):]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in recs:
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(recs)} launches, {tot:.3f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k[:64]:64s} n={a[0]:5d} {a[1]:9.3f} ms {100 * a[1] / tot:5.1f}%  {1e3 * a[1] / a[0]:8.1f} us/launch")
