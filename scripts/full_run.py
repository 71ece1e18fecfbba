"""configs[3] per-GPU share, complete: box 12, promoter 3, R replicas x 2000 chunks x 200 steps from distinct initial
conditions, thresholds 10..100 round-robin, fused measurements; then the 5-run contact tables on the device.
usage: python scripts/full_run.py [R] [chunks] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E, protocol as P

R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NCH = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/full_run.json"
t0 = time.time()
datas = [datafile.make_shoebox(12, 3, False, seed=k) for k in range(R)]
thr = np.array([10 * (1 + k % 10) for k in range(R)], np.int32)
eng = E.ShoeboxBatch(datas, n_replicas=R, seeds=np.arange(R, dtype=np.uint64) + 1, thresholds=thr, activations=30)
eng.set_schedule(NCH, observe=True)
t_setup = time.time() - t0
t0 = time.time()
eng.run(0, NCH)
t_run = time.time() - t0
status = eng.status()
frames = eng.frames()
c = eng.counters()
ok = status == 0
rows = E.aggregate_grouped(frames[ok][: (int(ok.sum()) // 5) * 5]) if ok.sum() >= 5 else np.zeros((0, 4))
res = {
    "workload": f"box 12, promoter 3, {R} replicas x {NCH} chunks x 200 steps, thresholds 10..100, -a 30",
    "replica_steps": int(c["steps"].sum()), "wall_s": t_run, "replica_steps_per_s": float(c["steps"].sum() / t_run),
    "setup_s": t_setup, "failed_replicas": int((~ok).sum()), "status_bits": [int(v) for v in np.unique(status)],
    "builds_per_step": float(c["builds"].sum() / max(c["steps"].sum(), 1)), "dangerous_per_step": float(c["dangerous"].sum() / max(c["steps"].sum(), 1)),
    "pairs_in_cut_per_eval_last": None,
    "active_fraction_by_threshold": {int(t): float(np.mean(frames[ok & (thr == t)][:, 150:, 8] == 2)) for t in np.unique(thr)},
    "mean_contact_pct": float(np.nanmean(rows[:, 2])) if len(rows) else None,
    "mean_S5PInt": float(np.nanmean(rows[:, 0])) if len(rows) else None,
    "groups": int(len(rows)),
    "final_rnp_fraction": float(np.mean(eng.state()[2] == 5) / np.mean(np.isin(eng.state()[2], (4, 5)))),
}
print(json.dumps(res, indent=1))
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
eng.close()
