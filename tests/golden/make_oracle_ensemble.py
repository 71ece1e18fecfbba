"""Generates tests/golden/oracle_ensemble_box11.npz: per-chunk ensemble statistics of the CPU oracle
(box 11, promoter 3, threshold 20, activation 100/1000) over N replicas x 400 chunks.
Run here (CPU): python tests/golden/make_oracle_ensemble.py <first_replica> <n_replicas> <out.npz>
then merge with --merge.  Used by tests/test_gpu_ensemble.py as the statistical reference.
ESB_ENSEMBLE_CHUNKS=2000 makes the full-length fixture oracle_ensemble_box11_2000.npz (production run length, 90 % RNP).
(Regenerated with the oracle's `neigh_modify delay 10` rule: 8 x `make_oracle_ensemble.py 6k 6 part_k.npz`, then
`--merge oracle_ensemble_box11.npz part_*.npz`.)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from enhancershoebox_b200 import datafile, ljsoft, protocol as P
from oracle import esb_oracle as O, observe as OB

NCH, THR, ACT = int(os.environ.get('ESB_ENSEMBLE_CHUNKS', '400')), 20, 100

def run(first, count, out):
    pt = P.pair_tables('genes_stages')
    secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
    lookup = [O.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
    plans = P.chunk_schedule('genes_stages', n_chunks=NCH)
    frames = np.zeros((count, NCH, 9)); ke = np.zeros((count, NCH)); nrnp = np.zeros((count, NCH), int); status = np.zeros(count, int)
    for r in range(count):
        rep = first + r
        d = datafile.make_shoebox(11, 3, False, seed=rep)
        o = O.Oracle(d, pt, lookup, seed=1000 + rep)
        lay = OB.Layout.from_data(d)
        gm = OB.GeneMachine(lay, threshold=THR, p_activation=ACT / 1000, seed=1000 + rep)
        exp = OB.expected_rnp_table(NCH, lay.n_rbprnp, 2000)
        types = d.types.copy()
        for pl in plans:
            st = o.run_chunk(pl)
            if st:
                status[r] = st
                break
            x, v, f, t = o.state()
            m = OB.measure(x, lay)
            s = gm.step(pl.index, types, m['n_prom'], int(exp[pl.index]))
            o.set_types(types)
            frames[r, pl.index] = OB.frame_record(pl.index, m, s)
            ke[r, pl.index] = 0.5 * (v ** 2).sum() / d.n_atoms
            nrnp[r, pl.index] = int((types == 5).sum())
        print('replica', rep, 'done', flush=True)
    np.savez_compressed(out, frames=frames, ke=ke, nrnp=nrnp, status=status)

if __name__ == '__main__':
    if sys.argv[1] == '--merge':
        parts = [np.load(p) for p in sys.argv[3:]]
        fr = np.concatenate([p['frames'] for p in parts])
        cat = lambda k: np.concatenate([p[k] for p in parts])
        # compact form read by tests/test_gpu_ensemble.py (only the columns the test compares)
        np.savez_compressed(sys.argv[2], d_rp=fr[:, :, 4].astype(np.float32), d_rg=fr[:, :, 5].astype(np.float32),
                            n_prom=fr[:, :, 6].astype(np.int16), n_gene=fr[:, :, 7].astype(np.int16), state=fr[:, :, 8].astype(np.int8),
                            ke=cat('ke').astype(np.float32), nrnp=cat('nrnp').astype(np.int16), status=cat('status'),
                            meta=np.array([NCH, THR, ACT]))
    else:
        run(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])
