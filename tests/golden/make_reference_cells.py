"""Golden ensemble statistics of the reference's OWN production sweep (box 11, 330 cells x 100 repeats, run with
LAMMPS by the reference's authors): VisitorGeneMaps/summary_contact_all_Thresholds10-100.txt, one row per repeat
(dist_genestages_all.py:66-85).  This script condenses that table into tests/golden/reference_cells_box11.json:

* `cells`: for every (Promoter, Threshold, Activation) cell the number of repeats, mean and standard deviation of
  S5PInt, S2PInt and Contact over the repeats, the number of repeats with an activation (DistActivation not NaN) and
  the mean / sd of DistActivation over those;
* `rows`: the 100 per-repeat rows of the five cells BASELINE.md quotes, for distribution-level tests.

Run here (needs /root/reference): python tests/golden/make_reference_cells.py"""
import json
import os
import sys

import numpy as np
import pandas as pd

SRC = "/root/reference/VisitorGeneMaps/summary_contact_all_Thresholds10-100.txt"
QUOTED = [(3, 80, 30), (1, 80, 30), (2, 50, 10), (3, 10, 100), (3, 100, 1)]


def main():
    df = pd.read_csv(SRC, index_col=0)
    cells, rows = [], {}
    for (p, t, a), g in df.groupby(["Promoter", "Threshold", "Activation"]):
        d = g["DistActivation"].to_numpy(float)
        act = d[~np.isnan(d)]
        cells.append(dict(promoter=int(p), threshold=int(t), activation=int(a), n=int(len(g)),
                          s5p=[float(g["S5PInt"].mean()), float(g["S5PInt"].std(ddof=1))],
                          s2p=[float(g["S2PInt"].mean()), float(g["S2PInt"].std(ddof=1))],
                          contact=[float(g["Contact"].mean()), float(g["Contact"].std(ddof=1))],
                          n_act=int(len(act)),
                          dist=[float(act.mean()) if len(act) else None, float(act.std(ddof=1)) if len(act) > 1 else None]))
        if (int(p), int(t), int(a)) in QUOTED:
            rows[f"{int(p)},{int(t)},{int(a)}"] = [[None if np.isnan(v) else float(v) for v in r]
                                                    for r in g[["S5PInt", "S2PInt", "Contact", "DistActivation"]].to_numpy(float)]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cells_box11.json")
    json.dump(dict(source="VisitorGeneMaps/summary_contact_all_Thresholds10-100.txt", n_rows=int(len(df)), cells=cells, rows=rows),
              open(out, "w"))
    print(out, len(cells), "cells", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
