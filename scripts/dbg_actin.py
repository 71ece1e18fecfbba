import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E, protocol as P
from oracle import actin as A
d = datafile.make_shoebox(11, 3, True, seed=11)
eng = E.ShoeboxBatch(d, n_replicas=1, driver="shoebox", condition="Control", seeds=[11], p_nucleation=0.05)
first, count = eng.actin.first, eng.actin.count
eng.set_schedule(22, observe=True); eng.run(0, 22)
st0 = eng.actin_state()
orc = A.ActinState(st0["free"][0], st0["filaments"][0], 11, 0.05)
x_pre, _, _ = eng.state()
pl = P.chunk_schedule("shoebox", "Control", n_chunks=40)[39]
pl = P.ChunkPlan(pl.index, pl.pair, None, False, False, False, 0)
eng.set_schedule(plans=[pl], observe=True)
eng.run(0, 1)
x_post, _, _ = eng.state()
st = eng.actin_state()
xo = x_pre[0].astype(np.float32).copy()
n = orc.events(39, xo, P.actin_flags(39))
print('oracle draws', n, 'free', len(orc.free_actin), 'fils', [len(f) for f in orc.filament_details])
print('device free', len(st["free"][0]), 'fils', [len(f) for f in st["filaments"][0]])
print('oracle fils', orc.filament_details[:6]); print('device fils', st["filaments"][0][:6])
print('moved device', np.flatnonzero(np.any(x_post[0].astype(np.float32) != x_pre[0].astype(np.float32), axis=1))[:20])
print('moved oracle', np.flatnonzero(np.any(xo != x_pre[0].astype(np.float32), axis=1))[:20])
print(eng.actin_frames(0,1)[0], eng.status())
