"""Device counterpart of scripts/oracle_actin_survival.py: fraction of actin replicas that died by chunk c."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E
R = int(sys.argv[1]) if len(sys.argv) > 1 else 192
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 400
pn = float(sys.argv[3]) if len(sys.argv) > 3 else 0.001
datas = [datafile.make_shoebox(11, 3, True, seed=100 + k) for k in range(R)]
eng = E.ShoeboxBatch(datas, driver="shoebox", condition="Control", seeds=np.arange(R) + 1, thresholds=80, activations=30, p_nucleation=pn)
eng.set_schedule(nch, observe=True)
for c0 in range(0, nch, 50):
    eng.run(c0, min(50, nch - c0))
    st = eng.status()
    print(f"after chunk {min(c0 + 50, nch)}: {int((st != 0).sum())}/{R} died  {dict(zip(*np.unique(st, return_counts=True)))}", flush=True)
stats, _ = eng.actin_frames()
ok = eng.status() == 0
print("alive: filaments", stats[ok, -1, 0].mean(), "sum len", stats[ok, -1, 1].mean(), "max len", stats[ok, -1, 5].mean(), "free", stats[ok, -1, 3].mean())
