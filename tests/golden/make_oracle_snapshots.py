import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np
from enhancershoebox_b200 import datafile, ljsoft, protocol as P
from oracle import esb_oracle as O, observe as OB
box=int(sys.argv[1]); nch=int(sys.argv[2]); out=sys.argv[3]
d = datafile.make_shoebox(box,3,False,seed=7)
pt = P.pair_tables('genes_stages')
secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
lookup = [O.build_lookup_table(secs[k], c) for (k,c) in pt.specs]
o = O.Oracle(d, pt, lookup, seed=4242)
lay = OB.Layout.from_data(d)
gm = OB.GeneMachine(lay, threshold=80, p_activation=30/1000, seed=4242)
exp = OB.expected_rnp_table(nch, lay.n_rbprnp, 2000)
plans = P.chunk_schedule('genes_stages', n_chunks=nch)
types = d.types.copy()
frames=[]; stats=[]; snaps={}
t0=time.time()
for pl in plans:
    o.reset_counters()
    st = o.run_chunk(pl)
    if st: print('FAILED status',st,'chunk',pl.index); break
    c=o.counters()
    x,v,f,t = o.state()
    m = OB.measure(x, lay)
    s = gm.step(pl.index, types, m['n_prom'], int(exp[pl.index]))
    o.set_types(types)
    frames.append(OB.frame_record(pl.index, m, s))
    stats.append([pl.index, c['pairs_in_cut']/c['evals'], c['pairs_listed']/c['evals'], c['builds'], int((types==5).sum())])
    if pl.index in (9,14,20,150,500,1000,1500,1999):
        snaps[f'x{pl.index}']=x; snaps[f'v{pl.index}']=v; snaps[f't{pl.index}']=types.copy()
    if pl.index%50==0:
        print(pl.index, '%.0fs'%(time.time()-t0), stats[-1], frames[-1][4:], flush=True)
np.savez(out, frames=np.array(frames,dtype=float), stats=np.array(stats), **snaps)
print('done', time.time()-t0)
