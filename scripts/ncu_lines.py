"""Aggregate an ncu source page (cuda,sass CSV) per CUDA source line: samples, instructions, top stall reasons.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python scripts/ncu_lines.py src.csv [N]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
agg = {}
stall = collections.defaultdict(collections.Counter)
tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != '-': continue
    try: n = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[4:], r[4:]))
    smp = float(d['# Samples'] or 0); ins = float(d['Instructions Executed'] or 0); thr = float(d['Thread Instructions Executed'] or 0)
    a = agg.setdefault((cur, n), [0.0, 0.0, 0.0, r[1]])
    a[0] += smp; a[1] += ins; a[2] += thr
    tot_s += smp; tot_i += ins
    for k in hdr:
        if k.startswith('stall_') and 'Not Issued' not in k:
            try: stall[(cur, n)][k] += float(d[k] or 0)
            except ValueError: pass
print('total samples', tot_s, 'total warp instructions', tot_i)
for (f, n), (s, i, t, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ', '.join(f'{k[6:]}:{v:.0f}' for k, v in stall[(f, n)].most_common(3))
    print(f'{f[:14]:14s}:{n:4d} smp {100*s/tot_s:5.2f}% inst {100*i/tot_i:5.2f}% thr/inst {t/max(i,1):4.1f}  {str(src).strip()[:64]:64s} | {top}')
