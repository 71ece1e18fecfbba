import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E
R = 64
datas = [datafile.make_shoebox(11, 3, True, seed=100 + k) for k in range(R)]
eng = E.ShoeboxBatch(datas, driver="shoebox", condition="Control", seeds=np.arange(R) + 1, thresholds=80, activations=30, p_nucleation=0.001)
eng.set_schedule(80, observe=True)
first = eng.actin.first
prev = None
for c in range(0, 80):
    eng.run(c, 1)
    st = eng.status()
    x, v, t = eng.state()
    if np.any(st):
        r = int(np.flatnonzero(st)[0])
        print('chunk', c, 'replica', r, 'status', st[r], 'steps', eng.counters()['steps'][r])
        lo = np.array(datas[0].box_lo); hi = np.array(datas[0].box_hi)
        out = np.flatnonzero(np.any((x[r] <= lo) | (x[r] >= hi), axis=1))
        sp = np.linalg.norm(v[r], axis=1)
        print('outside', out, 'types', t[r][out], 'pos', x[r][out], 'speed', sp[out], 'max speed', sp.max(), int(sp.argmax()), t[r][sp.argmax()])
        a = eng.actin_state()
        xp, vp, tp, ap = prev
        for b in list(out) + [int(sp.argmax())]:
            print('bead', b, 'prev pos', xp[r][b], 'prev v', vp[r][b], 'now', x[r][b])
            if b >= first:
                k = b - first
                print('  links now', a['plus'][r][k], a['minus'][r][k], 'prev links', ap['plus'][r][k], ap['minus'][r][k])
                for nb in (ap['plus'][r][k], ap['minus'][r][k]):
                    if nb >= 0: print('   prev dist to', nb, np.linalg.norm(xp[r][b] - xp[r][nb]))
        fl = [len(f) for f in ap['filaments'][r]]
        print('prev filaments', fl)
        break
    prev = (x, v, t, eng.actin_state())
