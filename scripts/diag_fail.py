"""Which replica of the 2000-chunk oracle-ensemble test configuration fails, with which status bits, and when."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E
z = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'oracle_ensemble_box11_2000.npz'))
nch, thr, act = [int(v) for v in z["meta"]]
R = int(sys.argv[1]) if len(sys.argv) > 1 else 192
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 777
datas = [datafile.make_shoebox(11, 3, False, seed=10_000 + r) for r in range(R)]
eng = E.ShoeboxBatch(datas, seeds=np.arange(R) + seed0, thresholds=thr, activations=act)
eng.set_schedule(nch, observe=True)
eng.run()
st = eng.status(); c = eng.counters()
bad = np.nonzero(st)[0]
print('failed', len(bad), 'of', R, [(int(b), int(st[b]), int(c['steps'][b])) for b in bad], 'max pair force/256', c['fmax'].max() if 'fmax' in c else None, flush=True)
print({k: (float(np.min(v)), float(np.max(v))) for k, v in c.items() if k in ('steps', 'fmax', 'danger')})
