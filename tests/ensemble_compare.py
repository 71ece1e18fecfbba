"""Level 3 of the north-star checks against the REFERENCE'S OWN results: per-run rows (S5PInt, S2PInt, Contact,
DistActivation) of production-protocol ensembles compared with the reference's LAMMPS sweep
(VisitorGeneMaps/summary_contact_all_Thresholds10-100.txt, condensed by tests/golden/make_reference_cells.py).

Both samples are finite (100 repeats per cell in the reference), the per-run distributions are broad and bimodal, so
means are compared through z = (m_dev - m_ref) / sqrt(s_dev^2/n_dev + s_ref^2/n_ref), activation fractions through the
pooled two-sample binomial z."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cells_box11.json")
COLUMNS = ("S5PInt", "S2PInt", "Contact", "DistActivation")
S2P_QUANTUM = 100.0 / 1850.0          # one frame of 2000 - 150 (Development/filter_ser2p_points.py:38-41; dist_genestages_all.py)


def _arr(rows):
    return np.array([[np.nan if v is None else v for v in r] for r in rows], float).reshape(-1, 4)


def load_golden():
    with open(GOLDEN) as f:
        g = json.load(f)
    cells = {(c["promoter"], c["threshold"], c["activation"]): c for c in g["cells"]}
    rows = {tuple(int(v) for v in k.split(",")): _arr(r) for k, r in g["rows"].items()}
    return cells, rows


def compare_cell(dev_rows, ref):
    """dev_rows [n, 4] per-run rows of the device; ref = one entry of golden `cells`.  Returns a list of
    (name, device mean, device sd, n, reference mean, reference sd, n, z)."""
    a = _arr(dev_rows)
    out = []
    for c, key in ((0, "s5p"), (1, "s2p"), (2, "contact")):
        m, s = ref[key]
        x = a[:, c]
        z = (x.mean() - m) / np.sqrt(x.var(ddof=1) / len(x) + s * s / ref["n"])
        out.append((COLUMNS[c], float(x.mean()), float(x.std(ddof=1)), len(x), m, s, ref["n"], float(z)))
    d = a[:, 3][~np.isnan(a[:, 3])]
    if len(d) > 1 and ref["n_act"] > 1:
        m, s = ref["dist"]
        z = (d.mean() - m) / np.sqrt(d.var(ddof=1) / len(d) + s * s / ref["n_act"])
        out.append((COLUMNS[3], float(d.mean()), float(d.std(ddof=1)), len(d), m, s, ref["n_act"], float(z)))
    # fraction of repeats with at least one activation
    p1, p2 = len(d) / len(a), ref["n_act"] / ref["n"]
    pp = (len(d) + ref["n_act"]) / (len(a) + ref["n"])
    se = np.sqrt(max(pp * (1 - pp), 1e-12) * (1 / len(a) + 1 / ref["n"]))
    out.append(("ActivatingFraction", p1, float("nan"), len(a), p2, float("nan"), ref["n"], float((p1 - p2) / se)))
    return out


def s2p_is_quantised(dev_rows) -> bool:
    """S2PInt of a 2000-frame run is a whole number of frames out of 1850."""
    q = _arr(dev_rows)[:, 1] / S2P_QUANTUM
    return bool(np.allclose(q, np.round(q), atol=1e-6))


def compare_result_file(path):
    """{cell: comparison list} for a result file of scripts/reference_ensemble.py."""
    with open(path) as f:
        res = json.load(f)
    cells, _ = load_golden()
    return {k: compare_cell(res["cells"][k], cells[tuple(int(v) for v in k.split(","))]) for k in res["cells"]}, res


def table(comparisons) -> str:
    lines = ["| cell (P,T,A) | quantity | device mean ± sd (n) | reference mean ± sd (n) | z |", "|---|---|---|---|---|"]
    for cell, comp in comparisons.items():
        for name, m1, s1, n1, m2, s2, n2, z in comp:
            lines.append(f"| ({cell}) | {name} | {m1:.3f} ± {s1:.3f} ({n1}) | {m2:.3f} ± {s2:.3f} ({n2}) | {z:+.2f} |")
    return "\n".join(lines)


def promoter_means(res, cells):
    """Mean over the cells of one promoter length of the cell means of S5PInt, S2PInt, Contact -- the rows of BASELINE.md's
    "mean over all cells, by promoter" -- for the device result `res` and the reference, with the standard error of each
    (cells are independent ensembles) and z of the difference.  {promoter: [(name, dev, dev_se, ref, ref_se, z)]}"""
    out = {}
    for p in (1, 2, 3):
        dev = [_arr(v) for k, v in res["cells"].items() if k.startswith(f"{p},")]
        ref = [cells[tuple(int(q) for q in k.split(","))] for k in res["cells"] if k.startswith(f"{p},")]
        if not dev:
            continue
        rows = []
        for c, key in ((0, "s5p"), (1, "s2p"), (2, "contact")):
            dm = float(np.mean([d[:, c].mean() for d in dev]))
            rm = float(np.mean([r[key][0] for r in ref]))
            dse = float(np.sqrt(sum(d[:, c].var(ddof=1) / len(d) for d in dev)) / len(dev))
            rse = float(np.sqrt(sum(r[key][1] ** 2 / r["n"] for r in ref)) / len(ref))
            rows.append((COLUMNS[c], dm, dse, rm, rse, (dm - rm) / float(np.hypot(dse, rse))))
        out[p] = rows
    return out
