import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, protocol as P, engine as E
d = datafile.make_shoebox(11, 3, False, seed=7)
snap = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'snap_box11.npz'))
x = snap['x20'].astype(np.float64); v = snap['v20'].astype(np.float64); t = snap['t20'].astype(np.int32)
xs = np.stack([x, x, x]); xs[1, 700, 0] = 13.5
eng = E.ShoeboxBatch(d, n_replicas=3, seeds=[1, 1, 1])
eng.upload_state(xs, np.stack([v] * 3), np.stack([t] * 3))
xg, _, _ = eng.state(); print('uploaded x[1,700]', xg[1, 700], 'types', t[700])
pl = P.chunk_schedule("genes_stages", n_chunks=22)[20:]
eng.set_schedule(plans=pl, observe=True, n_steps=50)
eng.run()
print('status', eng.status(), 'counters', {k: v.tolist() for k, v in eng.counters().items() if not k.startswith('cyc')})
xg, _, _ = eng.state(); print('x[1,700] after', xg[1, 700], xg[0, 700])
f, e = eng.forces('B', with_post=True)
print('hook status', eng.status())
