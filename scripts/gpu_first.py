"""Exploratory first GPU run: force parity vs the oracle at a few states, a short run, timing."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, ljsoft, protocol as P, engine as E
from oracle import esb_oracle as O, observe as OB

O.build()
d = datafile.make_shoebox(11, 3, False, seed=7)
pt = P.pair_tables('genes_stages')
t0 = time.time()
lookup, cfs = E.lookup_tables(pt)
print('tables built', time.time() - t0, flush=True)
# product builder vs oracle builder
secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
lo = [O.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
print('table builder max |diff| f:', max(np.abs(a[1] - b[1]).max() for a, b in zip(lookup, lo)))

snap = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'snap_box11.npz'))
orc = O.Oracle(d, pt, lo, seed=4242)
eng = E.ShoeboxBatch(d, n_replicas=2, seeds=[4242, 4243])
for tag, pair in (('9', 'soft'), ('14', 'A'), ('20', 'B'), ('500', 'B'), ('1999', 'B')):
    x = snap['x' + tag].astype(np.float64); v = snap['v' + tag].astype(np.float64); t = snap['t' + tag]
    orc.set_state(t, x, v); orc.set_pair(pair)
    if pair == 'soft':
        orc.L.esbo_pair_soft(orc.h, 54.0, O._dp(orc.soft_cut))
    fo, eo, st = orc.forces()
    eng.upload_state(np.stack([x, x]), np.stack([v, v]), np.stack([t, t]))
    for arrays in (True, False):
        fg, eg = eng.forces(pair, soft_a=54.0, use_arrays=arrays)
        scale = np.abs(fo).max()
        err = np.abs(fg[0] - fo).max() / scale
        print(f'chunk {tag} pair {pair} arrays={arrays}: max|df|/max|f| = {err:.3e} (fmax {scale:.2f}) E gpu {eg[0]} E orc {eo}', flush=True)
    # with post-force (noise, walls, anchors)
    orc.set_seed(4242, 77)
    fo2, _, _ = orc.forces(with_post=True, with_noise=True)
    fg2, _ = eng.forces(pair, soft_a=54.0, with_post=True, eval_index=77)
    print('   post-force err', np.abs(fg2[0] - fo2).max() / np.abs(fo2).max(), 'status', eng.status())

# short trajectory from the IC, compare to the oracle with identical noise
R = 64
eng = E.ShoeboxBatch(d, n_replicas=R, seeds=np.arange(R) + 4242, thresholds=80, activations=30)
eng.set_schedule(40)
t0 = time.time(); eng.run(0, 1); print('chunk 0 time', time.time() - t0)
orc = O.Oracle(d, pt, lo, seed=4242)
plans = P.chunk_schedule('genes_stages', n_chunks=40)
orc.run_chunk(plans[0])
xo, vo, fo, to = orc.state()
xg, vg, tg = eng.state()
print('after chunk 0: max|dx| vs oracle (same noise):', np.abs(xg[0] - xo).max(), ' replica spread', np.abs(xg[0] - xg[1]).max(), 'r0', eng.bond_r0(), orc.bond_r0(1), 'step', eng.step)
t0 = time.time(); eng.run(1, 39); dt = time.time() - t0
print(f'39 chunks x {R} replicas: {dt:.3f}s -> {39*200*R/dt:.3e} replica-steps/s', 'status', np.bincount(eng.status()))
c = eng.counters()
print({k: (v[:3].tolist()) for k, v in c.items()})
fr = eng.frames()
print('frames[0][-3:]', fr[0][-3:])
# bigger batch timing at a late-time state
R = 444
x = snap['x1999'].astype(np.float64); v = snap['v1999'].astype(np.float64); t = snap['t1999']
eng2 = E.ShoeboxBatch(d, n_replicas=R, seeds=np.arange(R) + 1)
eng2.upload_state(x[None], v[None], t[None], replicate=True)
plans = [pl for pl in P.chunk_schedule('genes_stages', n_chunks=2000)][1990:2000]
eng2.set_schedule(plans=plans, observe=True)
eng2.run(0, 2)
eng2.reset_counters()
t0 = time.time(); eng2.run(2, 8); dt = time.time() - t0
c = eng2.counters()
print(f'late state: 8 chunks x {R}: {dt:.3f}s -> {8*200*R/dt:.3e} replica-steps/s; pairs/eval {c["pairs"].sum()/c["evals"].sum()/2:.0f} listed/eval {c["listed"].sum()/c["evals"].sum()/2:.0f} builds/chunk {c["builds"].mean()/8:.1f}', 'status', np.bincount(eng2.status()))
