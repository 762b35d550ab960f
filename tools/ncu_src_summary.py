#!/usr/bin/env python
"""Summarise `ncu --page source --print-source cuda,sass --csv` (argv[1]) per CUDA source line and `--page source --csv`
(argv[2], SASS) by active lanes: where a kernel's instructions and stall samples go."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0]); src = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name': continue
    if r[0].strip().isdigit() and hdr:
        n = len(hdr); ie = hdr.index('Instructions Executed'); sm = hdr.index('# Samples'); off = len(r) - n
        try: inst = int(r[ie + off]); samp = int(r[sm + off])
        except Exception: continue
        key = (cur, int(r[0])); agg[key][0] += inst; agg[key][1] += samp; src[key] = ','.join(r[1:2 + off])[:100]
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('warp instructions', tot, 'samples', ts)
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{k[0]:22s} {k[1]:5d} {100*v[0]/tot:5.1f}% inst {100*v[1]/max(ts,1):5.1f}% smp  {src[k]}")
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2]))); H = rows[1]
    ia = H.index('Instructions Executed'); ip = H.index('Predicated-On Thread Instructions Executed')
    bins = collections.Counter(); t = 0; tp = 0
    for r in rows[2:]:
        try: a = int(r[ia]); p = int(r[ip])
        except Exception: continue
        if not a: continue
        l = p / a; t += a; tp += p
        bins['0-4' if l < 4 else '4-8' if l < 8 else '8-16' if l < 16 else '16-24' if l < 24 else '24-30' if l < 30 else '30-32'] += a
    print('avg active lanes %.1f' % (tp / t), {b: round(100 * bins[b] / t, 1) for b in ['0-4', '4-8', '8-16', '16-24', '24-30', '30-32']})
