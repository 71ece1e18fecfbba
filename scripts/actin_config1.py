"""BASELINE.json configs[1] to completion: box-11 actin shoebox (-z 300, promoter 3), 64 seeds on one GPU, 2000 Python
timesteps x 200 MD steps with the polymerisation events on.  A repeat that dies (LAMMPS: lost atoms / particle on or
inside the wall -- README.md:203-205) is deleted and re-run with a fresh seed until it completes, which is what the
reference's launcher does (run_parallel_shoebox.sh:1-14 `task()`: rm -r run<r>; python run_single_shoebox.py ... until
the sentinel exists).  Records how many attempts each repeat needed and when the failed attempts died.
usage: python scripts/actin_config1.py [seeds] [chunks] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from enhancershoebox_b200 import datafile, engine as E

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NCH = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/r2_actin_config1.json"
pending = list(range(N))
attempts = np.zeros(N, int)
death_chunks, final = [], {}
t0 = time.time(); steps = 0; rounds = 0
while pending and rounds < 60:
    rounds += 1
    # the launcher keeps 8 processes busy; here every round tops the batch up to N replicas so that the GPU stays full:
    # repeats still pending get several fresh seeds at once, the first that completes is kept
    copies = max(1, N // len(pending))
    runs = [(r, k) for r in pending for k in range(copies)]
    seeds = np.array([1 + r + 1009 * (attempts[r] + k) for r, k in runs], dtype=np.uint64)
    datas = [datafile.make_shoebox(11, 3, True, seed=int(s)) for s in seeds]
    eng = E.ShoeboxBatch(datas, driver="shoebox", condition="Control", seeds=seeds, thresholds=80, activations=30, p_nucleation=0.001)
    eng.set_schedule(NCH, observe=True)
    eng.run()
    st = eng.status(); c = eng.counters(); steps += int(c["steps"].sum())
    stats, _ = eng.actin_frames()
    fr = eng.frames()
    eng.close()
    done = set()
    for q, (r, k) in enumerate(runs):
        if st[q] == 0 and r not in done:
            done.add(r)
            final[r] = dict(attempts=int(attempts[r] + k + 1), filaments=int(stats[q, -1, 0]), mean_len=float(stats[q, -1, 1] / max(stats[q, -1, 0], 1)),
                            max_len=int(stats[q, -1, 5]), free=int(stats[q, -1, 3]), active_frames=int((fr[q, :, 8] == 2).sum()))
        elif st[q] != 0:
            death_chunks.append(int(c["steps"][q] // 200))
    for r in pending:
        attempts[r] += copies
    pending = [r for r in pending if r not in done]
    print(f"round {rounds}: {len(runs)} runs, {len(done)} repeats completed, {len(pending)} pending, {time.time() - t0:.0f} s", flush=True)
att = np.array([final[r]["attempts"] for r in sorted(final)])
res = dict(workload=f"box 11 actin (-z 300, -n 0.001), promoter 3, {N} repeats x {NCH} chunks x 200 steps, events on", completed=len(final), pending=len(pending),
           wall_s=time.time() - t0, replica_steps=steps, rounds=rounds, attempts_mean=float(att.mean()) if len(att) else None, attempts_max=int(att.max()) if len(att) else None,
           survival_per_attempt=float(len(final) / max(len(final) + len(death_chunks), 1)),
           death_chunk_hist=np.histogram(death_chunks, bins=[0, 40, 100, 200, 400, 800, 1200, 1600, 2000])[0].tolist(), death_chunk_bins=[0, 40, 100, 200, 400, 800, 1200, 1600, 2000],
           final={str(r): final[r] for r in sorted(final)})
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "final"}))
