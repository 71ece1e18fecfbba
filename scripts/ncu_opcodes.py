"""Executed warp instructions per SASS opcode from an ncu source page (cuda,sass CSV), only the kernel with the most instructions.
usage: python scripts/ncu_opcodes.py src.csv [N]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; fn = None
per_fn = collections.defaultdict(lambda: collections.Counter())
seen = collections.defaultdict(set)
for r in rows:
    if len(r) >= 2 and r[0] == 'Function Name': fn = r[1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] != '': continue
    addr = r[2]
    if addr in seen[fn]: continue          # (the SASS of a function is repeated under every file it has lines in)
    seen[fn].add(addr)
    d = dict(zip(hdr[4:], r[4:]))
    try: ins = float(d['Instructions Executed'] or 0)
    except ValueError: continue
    op = r[3].strip().split()
    if not op: continue
    o = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
    per_fn[fn][o.split('.')[0]] += ins
fn = max(per_fn, key=lambda f: sum(per_fn[f].values()))
tot = sum(per_fn[fn].values())
print(fn[:100], 'total', tot)
for o, v in per_fn[fn].most_common(topn): print(f'{o:12s} {100*v/tot:6.2f}%')
