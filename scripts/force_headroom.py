"""Largest pair-force component per chunk over a batch of production runs (headroom of the fixed-point force
accumulators).  usage: python scripts/force_headroom.py [R] [chunks] [box]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E

R = int(sys.argv[1]) if len(sys.argv) > 1 else 296
NCH = int(sys.argv[2]) if len(sys.argv) > 2 else 60
box = int(sys.argv[3]) if len(sys.argv) > 3 else 11
datas = [datafile.make_shoebox(box, 3, False, seed=10_000 + r) for r in range(R)]
eng = E.ShoeboxBatch(datas, seeds=np.arange(R) + 777, thresholds=20, activations=100)
eng.set_schedule(max(NCH, 400), observe=True)
for c in range(NCH):
    eng.reset_counters()
    eng.run(c, 1)
    m = eng.counters()["max_pair_force"]
    print(f"chunk {c:4d} ({eng.plans[c].pair:4s}): max |pair force| over {R} replicas {m.max():10.1f}  median {np.median(m):8.1f}  99% {np.percentile(m, 99):9.1f}  status {np.bincount(eng.status())}", flush=True)
eng.reset_counters()
eng.run(NCH, 400 - NCH)
m = eng.counters()["max_pair_force"]
print(f"chunks {NCH}..399: max {m.max():.1f} median {np.median(m):.1f} 99% {np.percentile(m, 99):.1f} status {np.bincount(eng.status())}")
