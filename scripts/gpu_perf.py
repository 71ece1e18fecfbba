"""Timing at a late-time state (chunk 1999 snapshot of box 11): phase breakdown from the in-kernel cycle counters."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, protocol as P, engine as E
R = int(sys.argv[1]) if len(sys.argv) > 1 else 444
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
d = datafile.make_shoebox(11, 3, False, seed=7)
snap = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'snap_box11.npz'))
x = snap['x1999'].astype(np.float64); v = snap['v1999'].astype(np.float64); t = snap['t1999']
eng = E.ShoeboxBatch(d, n_replicas=R, seeds=np.arange(R) + 1)
eng.upload_state(x[None], v[None], t[None], replicate=True)
plans = P.chunk_schedule('genes_stages', n_chunks=2000)[2000 - nch - 2:]
eng.set_schedule(plans=plans, observe=True)
eng.run(0, 2)
eng.reset_counters()
t0 = time.time(); eng.run(2, nch); dt = time.time() - t0
c = eng.counters()
ev = c['evals'].sum()
print(f'R={R}: {nch} chunks in {dt:.3f}s -> {nch*200*R/dt:.3e} replica-steps/s; pairs/eval {c["pairs"].sum()/ev/2:.0f} listed/eval {c["listed"].sum()/ev/2:.0f} builds/chunk {c["builds"].mean()/nch:.1f} status {np.bincount(eng.status())}')
tot = (c['cyc_atom'] + c['cyc_build'] + c['cyc_force'] + c['cyc_other']).mean()
for k in ('cyc_atom', 'cyc_build', 'cyc_force', 'cyc_other'):
    per = c[k].mean()
    print(f'  {k:10s} {per/ (nch*200):10.0f} cycles/step  ({100*per/tot:5.1f}%)' + (f'   per build {c[k].mean()/c["builds"].mean():.0f}' if k == 'cyc_build' else ''))
