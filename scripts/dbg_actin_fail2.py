import sys, time, os, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E, ljsoft, protocol as P
from oracle import esb_oracle as O
R = 64
datas = [datafile.make_shoebox(11, 3, True, seed=100 + k) for k in range(R)]
eng = E.ShoeboxBatch(datas, driver="shoebox", condition="Control", seeds=np.arange(R) + 1, thresholds=80, activations=30, p_nucleation=0.001)
eng.set_schedule(80, observe=True)
first = eng.actin.first
eng.run(0, 55)
r = 46
x, v, t = eng.state()
a = eng.actin_state()
d = datas[r]
bonds = [list(b) for b in d.bonds if b[1] < first]; angles = [list(g) for g in d.angles if g[1] < first]
for f in a['filaments'][r]:
    bonds += [[2, p, q] for p, q in zip(f[:-1], f[1:])]
    angles += [[2, p, q, s] for p, q, s in zip(f[:-2], f[1:-1], f[2:])]
dr = dataclasses.replace(d, bonds=np.array(bonds, np.int32), angles=np.array(angles, np.int32))
pt = P.pair_tables("shoebox", "Control")
secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
lookup = [O.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
o = O.Oracle(dr, pt, lookup, seed=r + 1)
o.set_state(t[r], x[r], v[r])
o.L.esbo_set_bond_coeff(o.h, 1, P.BOND_K, float(eng.bond_r0()[1]))
o.set_seed(r + 1, 55 * 201)
plan = P.chunk_schedule("shoebox", "Control", n_chunks=56)[55]
o.apply_plan(plan)
for n in (50, 50, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10):
    st = o.L.esbo_run(o.h, n, -1, -1) if False else None
    break
# run in pieces with the ramp of the whole 200-step run: use run_chunk-like start/stop
first_step = o.step
for k in range(20):
    st = o.L.esbo_run(o.h, 10, first_step, first_step + 200)
    xo, vo, fo, to = o.state()
    sp = np.linalg.norm(vo, axis=1)
    print('oracle step', o.step - first_step, 'status', st, 'max speed %.1f at %d' % (sp.max(), sp.argmax()), 'r0_2 %.4f' % o.bond_r0(2),
          'min actin bond %.4f' % min(np.linalg.norm(xo[p] - xo[q]) for (_, p, q) in bonds if p >= first))
    if st: break
