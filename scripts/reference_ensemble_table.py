"""Markdown table of a scripts/reference_ensemble.py result against the reference's LAMMPS sweep
(tests/golden/reference_cells_box11.json).  usage: python scripts/reference_ensemble_table.py result.json [--summary]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import ensemble_compare as EC

comp, res = EC.compare_result_file(sys.argv[1])
zs = np.array([q[-1] for c in comp.values() for q in c if np.isfinite(q[-1])])
print(f"box {res['box']}, {res['chunks']} chunks x 200 steps, {res['repeats']} repeats per cell, {len(comp)} cells, "
      f"{res['replica_steps']} replica-steps in {res['wall_s']:.1f} s on one B200, {res['failed']} failed runs\n")
print(f"{len(zs)} comparisons: mean z {zs.mean():+.3f}, rms z {np.sqrt((zs ** 2).mean()):.3f}, max |z| {np.abs(zs).max():.2f}, "
      f"sum z^2 {float((zs ** 2).sum()):.1f}\n")
if "--summary" in sys.argv:
    by = {}
    for c in comp.values():
        for q in c:
            if np.isfinite(q[-1]):
                by.setdefault(q[0], []).append(q[-1])
    print("| quantity | cells | mean z | rms z | max abs z |\n|---|---|---|---|---|")
    for k, v in by.items():
        v = np.array(v)
        print(f"| {k} | {len(v)} | {v.mean():+.3f} | {np.sqrt((v ** 2).mean()):.3f} | {np.abs(v).max():.2f} |")
    cells, _ = EC.load_golden()
    print("\n| promoter | quantity | device mean over cells +- se | reference mean over cells +- se | z |\n|---|---|---|---|---|")
    for p, rows in EC.promoter_means(res, cells).items():
        for name, dm, dse, rm, rse, z in rows:
            print(f"| {p} | {name} | {dm:.2f} +- {dse:.2f} | {rm:.2f} +- {rse:.2f} | {z:+.2f} |")
else:
    print(EC.table(comp))
