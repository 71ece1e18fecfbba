"""How often does the reference's actin protocol (fix s3 re-ramping the actin bond r0 from 0.033 in every run, K_angle = 33,
dt = 0.005) blow up in the fp64 oracle?  Replicas of the box-11 actin shoebox with the polymerisation events on
(oracle/actin.py), chunk by chunk; prints the chunk at which each replica died (status != 0) or 'ok'.
usage: python scripts/oracle_actin_survival.py <first_seed> <n_replicas> <chunks> [p_nucleation]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, ljsoft, protocol as P
from oracle import esb_oracle as O, actin as A

first_seed, n_rep, nch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
p_nuc = float(sys.argv[4]) if len(sys.argv) > 4 else 0.001
O.build()
pt = P.pair_tables("shoebox", "Control")
secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
lookup = [O.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
plans = P.chunk_schedule("shoebox", "Control", n_chunks=nch)
for rep in range(first_seed, first_seed + n_rep):
    d = datafile.make_shoebox(11, 3, True, seed=100 + rep)
    ids = np.flatnonzero(d.types == P.ACTIN_TYPE)
    first, count = int(ids[0]), len(ids)
    o = O.Oracle(d, pt, lookup, seed=rep + 1)
    st = A.ActinState(list(range(first + 3, first + count)), [[first, first + 1, first + 2]], seed=rep + 1, p_nucleation=p_nuc)
    static_b = [list(b) for b in d.bonds if b[1] < first]
    static_a = [list(g) for g in d.angles if g[1] < first]
    died, vmax = None, 0.0
    for pl in plans:
        if o.run_chunk(pl) != 0:
            died = pl.index
            break
        flags = P.actin_flags(pl.index)
        x, v, f, t = o.state()
        vmax = max(vmax, float(np.abs(v).max()))
        if flags & 1:
            st.events(pl.index, x, flags)
            o.set_state(t, x, v)
            b = np.array(static_b + [[2, p, q] for (p, q) in sorted(st.bonds)], np.int32)
            a = np.array(static_a + [[2, p, q, r] for (p, q, r) in sorted(st.angles)], np.int32)
            o.L.esbo_set_topology(o.h, b.shape[0], O._ip(b), a.shape[0], O._ip(a))
    lens = [len(fl) for fl in st.filament_details]
    print(f"replica {rep}: {'died at chunk %d' % died if died is not None else 'ok'}; filaments {len(lens)} (longest {max(lens) if lens else 0}), free {len(st.free_actin)}, max |v| {vmax:.1f}", flush=True)
