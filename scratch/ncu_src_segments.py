"""Aggregate an `ncu --page source --csv` export of one kernel by barrier-delimited segment (= phase of the FFT kernels):
samples, instructions, shared-memory wavefronts (ideal / excess), global L1 tag requests, top stall reasons.
usage: python scratch/ncu_src_segments.py file.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
seg = collections.OrderedDict()
cur = 0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot_s = sum(f(r, '# Samples') for r in rows[2:] if len(r) == len(hdr))
for r in rows[2:]:
    if len(r) != len(hdr): continue
    d = seg.setdefault(cur, collections.Counter())
    d['samples'] += f(r, '# Samples'); d['inst'] += f(r, 'Instructions Executed')
    d['sh_wave'] += f(r, 'L1 Wavefronts Shared'); d['sh_ideal'] += f(r, 'L1 Wavefronts Shared Ideal')
    d['g_tag'] += f(r, 'L1 Tag Requests Global'); d['n'] += 1
    for s in stalls: d[s] += f(r, s)
    if 'BAR.SYNC' in r[ix['Source']]: cur += 1
print(f"total samples {tot_s:.0f}")
for k, d in seg.items():
    top = sorted(((d[s], s) for s in stalls), reverse=True)[:4]
    print(f"seg {k:2d} n={d['n']:5.0f} samples {100*d['samples']/tot_s:5.1f}% inst {d['inst']/1e6:8.2f}M sh_wave {d['sh_wave']/1e6:7.2f}M (ideal {d['sh_ideal']/1e6:7.2f}M) g_tag {d['g_tag']/1e6:7.2f}M  " + ' '.join(f"{s[6:]}={100*v/max(d['samples'],1):.0f}%" for v, s in top))
