"""A small case for compute-sanitizer (one tool per gpurun call): the full-batch kernel (segmented and plain), the
32-warp kernel, the cluster kernel and the parity hook on a few box-9 replicas, two 12-step chunks each with the fused
observation, from a tabulated-phase frame.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E, protocol as P

z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "snap_box9.npz"))
d = datafile.make_shoebox(9, 3, False, seed=7)
x, v, t = z["x500"].astype(np.float64), z["v500"].astype(np.float64), z["t500"].astype(np.int32)
plans = [P.ChunkPlan(p.index, p.pair, p.ramp, p.adapt_soft, p.adapt_bond1, p.adapt_bond2, 12) for p in P.chunk_schedule("genes_stages", n_chunks=502)[500:]]
out = {}
for name, R, env in (("cluster8", 2, {}), ("wide", 3, {"ESB_CLUSTER_SIZE": "1"}), ("full", 3, {"ESB_WIDE": "0", "ESB_CLUSTER_SIZE": "1"}),
                     ("full_seg4", 3, {"ESB_WIDE": "0", "ESB_CLUSTER_SIZE": "1", "ESB_SEG": "4"})):
    for k in ("ESB_WIDE", "ESB_CLUSTER_SIZE", "ESB_SEG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = E.ShoeboxBatch(d, n_replicas=R, seeds=[7] * R)
    eng.upload_state(x[None], v[None], t[None], replicate=True)
    eng.set_schedule(plans=plans, observe=True)
    eng.run()
    f, e = eng.forces("B")
    xs, vs, ts = eng.state()
    assert np.all(eng.status() == 0)
    out[name] = (xs[0], vs[0], f[0])
    eng.close()
    print(name, "ok", flush=True)
ref = out["full"]
for name, o in out.items():
    assert all(np.array_equal(a, b) for a, b in zip(o, ref)), name
print("all variants bitwise identical")
