"""Survival of the configs[1] actin batch (64 repeats, events on) over the first chunks, with the status bits of the dead.
usage: [ESB_LIB=...] [ESB_CLUSTER_SIZE=1 ESB_WIDE=0] python scripts/diag_actin.py [R] [chunks] [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 400
tag = sys.argv[3] if len(sys.argv) > 3 else ""
datas = [datafile.make_shoebox(11, 3, True, seed=100 + k) for k in range(R)]
eng = E.ShoeboxBatch(datas, driver="shoebox", condition="Control", seeds=np.arange(R) + 1, thresholds=80, activations=30, p_nucleation=0.001)
eng.set_schedule(nch, observe=True)
alive = []
for c0 in range(0, nch, 50):
    eng.run(c0, min(50, nch - c0))
    alive.append(int((eng.status() == 0).sum()))
st = eng.status(); c = eng.counters()
print(f'{tag:10s} R={R} alive after every 50 chunks {alive}; status bits of the dead {dict(zip(*np.unique(st[st != 0], return_counts=True)))}; max pair force/256 {c["fmax"].max() if "fmax" in c else "?"}', flush=True)
