"""configs[1]: box-11 actin shoebox, 64 seeds on one GPU, polymerisation events on (from Python timestep 40).
usage: [ESB_WIDE=0|1] python scripts/perf_actin.py [R] [chunks] [p_nucleation]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 80
pn = float(sys.argv[3]) if len(sys.argv) > 3 else 0.001
datas = [datafile.make_shoebox(11, 3, True, seed=100 + k) for k in range(R)]
eng = E.ShoeboxBatch(datas, driver="shoebox", condition="Control", seeds=np.arange(R) + 1, thresholds=80, activations=30, p_nucleation=pn)
eng.set_schedule(nch, observe=True)
eng.run(0, 20)
eng.reset_counters()
t0 = time.time(); eng.run(20, nch - 20); dt = time.time() - t0
c = eng.counters()
st, _ = eng.actin_frames()
print(f'wide={os.environ.get("ESB_WIDE", "auto")} R={R} chunks 20..{nch}: {(nch-20)*200*R/dt:.3e} replica-steps/s ({dt:.2f} s); '
      f'filaments/replica at the end {st[:, -1, 0].mean():.1f}, mean length {st[:, -1, 1].sum() / max(st[:, -1, 0].sum(), 1):.2f}, '
      f'free actin {st[:, -1, 3].mean():.1f}; ok {int((eng.status() == 0).sum())}/{R} status {dict(zip(*np.unique(eng.status(), return_counts=True)))} steps {c["steps"].min()}..{c["steps"].max()}', flush=True)
