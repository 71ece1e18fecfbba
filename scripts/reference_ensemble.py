"""Production-protocol ensembles on the device for comparison with the reference's own LAMMPS sweep
(VisitorGeneMaps/summary_contact_all_Thresholds10-100.txt, box 11, 100 repeats per cell).

    python scripts/reference_ensemble.py [--cells quoted|all] [--repeats 100] [--chunks 2000] [--out gpurun_out/ref_ensemble.json]

Every repeat is what VisitorGeneMaps/run_single_GENES_STAGES.py -b 11 -p P -x T -a A does (fresh make_input_file
initial condition, soft phase, table switch, 2000 Python timesteps of 200 MD steps, gene state machine, RBP->RNP
ramp) and the per-run row is dist_genestages_all.py's (S5PInt, S2PInt, Contact, DistActivation), computed on the
device.  Writes the per-run rows per cell; tests/test_reference_ensemble.py compares them with the golden table."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from enhancershoebox_b200 import datafile, engine as E

QUOTED = [(3, 80, 30), (1, 80, 30), (2, 50, 10), (3, 10, 100), (3, 100, 1)]
THRESHOLDS = [10, 20, 30, 40, 50, 60, 70, 75, 80, 90, 100]
ACTIVATIONS = [1, 5, 10, 15, 20, 25, 30, 50, 75, 100]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", default="quoted")
    ap.add_argument("--repeats", type=int, default=100)
    ap.add_argument("--chunks", type=int, default=2000)
    ap.add_argument("--box", type=int, default=11)
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--seed-base", type=int, default=7_000_000)
    ap.add_argument("--ic", default="fresh", choices=["fresh", "bundled"],
                    help="fresh: a new make_input_file configuration per repeat; bundled: every repeat starts from the reference's one "
                         "bundled data file of that promoter length, as the reference's launcher does (tests/golden/bundled_ic_bs11.npz)")
    ap.add_argument("--out", default="gpurun_out/ref_ensemble.json")
    args = ap.parse_args()
    cells = QUOTED if args.cells == "quoted" else [(p, t, a) for p in (1, 2, 3) for a in ACTIVATIONS for t in THRESHOLDS]
    res = {"box": args.box, "chunks": args.chunks, "repeats": args.repeats, "ic": args.ic, "cells": {}, "failed": 0}
    t_all = time.time()
    steps = 0
    for p in (1, 2, 3):
        runs = [(c, k) for c in cells if c[0] == p for k in range(args.repeats)]
        for lo in range(0, len(runs), args.batch):
            part = runs[lo:lo + args.batch]
            seeds = np.array([args.seed_base + 100_003 * (c[1] * 1000 + c[2]) + 1_000_000_007 * p + k for c, k in part], dtype=np.uint64)
            if args.ic == "bundled":
                import dataclasses
                z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bundled_ic_bs11.npz"))
                d0 = datafile.make_shoebox(11, p, False, seed=0)
                datas = dataclasses.replace(d0, x=z[f"x{p}"].astype(np.float64), types=z[f"t{p}"].astype(np.int32))
            else:
                datas = [datafile.make_shoebox(args.box, p, False, seed=int(s)) for s in seeds]
            eng = E.ShoeboxBatch(datas, n_replicas=len(part), seeds=seeds, thresholds=[c[1] for c, _ in part], activations=[c[2] for c, _ in part])
            eng.set_schedule(args.chunks, observe=True)
            t0 = time.time()
            eng.run()
            dt = time.time() - t0
            st = eng.status()
            rows = E.aggregate_all(eng.frames())
            cn = eng.counters()
            steps += int(cn["steps"].sum())
            eng.close()
            print(f"promoter {p}: {len(part)} runs in {dt:.1f} s, {int((st != 0).sum())} failed", flush=True)
            for (c, k), r, s in zip(part, rows, st):
                if s != 0:
                    res["failed"] += 1
                    continue
                res["cells"].setdefault(f"{c[0]},{c[1]},{c[2]}", []).append([None if np.isnan(v) else float(v) for v in r])
    res["wall_s"] = time.time() - t_all
    res["replica_steps"] = steps
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"))
    print(f"{len(res['cells'])} cells, {steps} replica-steps, {res['wall_s']:.1f} s, {res['failed']} failed runs -> {args.out}")


if __name__ == "__main__":
    sys.exit(main())
