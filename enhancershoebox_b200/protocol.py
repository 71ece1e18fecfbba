"""The reference drivers' LAMMPS command stream, restated as data.

Two drivers are covered:

* ``genes_stages`` -- ``VisitorGeneMaps/run_single_GENES_STAGES.py`` (no actin; the sweep driver)
* ``shoebox``      -- ``run_single_shoebox.py`` (actin-capable)

`soft_cutoffs` replays the ``pair_style soft`` / ``pair_coeff`` block, `table_map` replays the
``pair_style table`` / ``pair_coeff`` block(s) with LAMMPS' wildcard and override rules
([LAMMPS-ext] ``utils::bounds``: "a*b" is a range, "*" is all; only pairs with I <= J are set;
a later pair_coeff overrides an earlier one), and `chunk_schedule` lays out what each Python
timestep ("chunk": one ``run 200``) does.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .datafile import N_ATOM_TYPES

MAXT = 12  # table side: atom types are 1..11, row/col 0 unused

# timescales (run_single_shoebox.py:372-375, 456, 469-475, 517-526)
N_RUNS = 2000
T_RUN = 200
DT = 0.005
DELT = 0.03 * T_RUN
T_SOFT = 10
T_SER5P_ON = 15
T_INDUCTION_ON = 150
ACTIVATION_DELAY = 2
T_ACTIVE_DURATION = 50
F_RNP_FINAL = 0.9
SIG_CHROMATIN_NM = 60
POL2_RELEASE_RADIUS = 3
CLUSTER_R = np.linspace(0.65, 2.5, 50)  # run_single_shoebox.py:425

# neighbor 0.3 multi; neigh_modify every 1 delay 10 check yes (run_single_shoebox.py:532-533)
NEIGH_SKIN = 0.3
NEIGH_DELAY = 10
# fixes (run_single_shoebox.py:643-647)
LANGEVIN_T = 1.0
LANGEVIN_DAMP = 0.5
WALL_EPS = 100.0
WALL_CUT = 3.0
# bonds / angles (run_single_shoebox.py:582-584, 631-639)
BOND_K = 90.0
BOND_R0 = {1: 1.0, 2: 0.3}
BOND_R0_RAMP = {1: (0.033, 1.0), 2: (0.033, 0.3)}
ANGLE_K = {1: 100 / SIG_CHROMATIN_NM, 2: 33.0}
SOFT_A_RAMP = (0.0, 60.0)
TABLE_LENGTH = 30000  # pair_style table lookup 30000

# actin dynamics (run_single_shoebox.py:398-423, 456-466, 489-496)
ACTIN_TYPE = 3
FILAMENT_LENGTH = 3
T_POLYMERIZATION_ON = 40
T_ON_PLUS = 1
T_ON_MINUS = 40
T_OFF_MINUS = 5
T_OFF_PLUS = 2 * N_RUNS                 # run_single_shoebox.py:496
P_ON = 0.3
SIG_ACTIN = 0.3
W_SA = 0.2
FIL_END_RADIUS = 3
MONOMER_RADIUS = 2
STATIC_ACTIN_CONDITIONS = ("LatB", "SwinA_nopol")            # t_polymerization_on = 2*NRuns (:459-460)


def actin_flags(i: int, condition: str = "Control") -> int:
    """Which actin events Python timestep i performs (run_single_shoebox.py:969, 1064, 1101, 1168):
    1 = nucleation + plus-end polymerisation, 2 = minus-end polymerisation, 4 = minus-end depolymerisation,
    8 = plus-end depolymerisation (i % t_off_plus == 0 with t_off_plus = 2 NRuns, :496, 1220-1250: once, at i = 4000, in
    the 5000-chunk JQ-1 / Flavopiridol runs)."""
    if condition in STATIC_ACTIN_CONDITIONS or (i + 1) < T_POLYMERIZATION_ON:
        return 0
    return (1 | (2 if i % (T_ON_PLUS * T_ON_MINUS) == 0 else 0) | (4 if i % T_OFF_MINUS == 0 else 0)
            | (8 if i % T_OFF_PLUS == 0 else 0))


# synthetic microscopy (run_single_shoebox.py:152-213, 1498-1502)
MICROSCOPY_EVERY = 10
MICROSCOPY_EDGES = (np.linspace(-11, 11, 22), np.linspace(-9, 9, 18), np.linspace(-7, 7, 14))
MICROSCOPY_CHANNELS = (("Ser5P", (8,)), ("Ser2P", (7, 11)), ("Chromatin", (1, 2, 6, 7, 9, 10, 11)), ("Gene", (2, 6, 7, 10, 11)), ("Actin", (3,)))
MICROSCOPY_SIGMA = 1


TREATMENT_DURATION = {  # run_single_shoebox.py:379-385
    "Hexanediol": int(3 * 60 / 0.6), "JQ-1": int(30 * 60 / 0.6), "Flavopiridol": int(30 * 60 / 0.6),
}


def _bounds(spec, n=N_ATOM_TYPES):
    spec = str(spec)
    if spec == "*":
        return 1, n
    if "*" in spec:
        lo, hi = spec.split("*")
        return (int(lo) if lo else 1), (int(hi) if hi else n)
    return int(spec), int(spec)


def _replay(stream, n=N_ATOM_TYPES):
    """Apply (I, J, value) commands in order; return dict[(i,j)] for i <= j."""
    out = {}
    for I, J, val in stream:
        ilo, ihi = _bounds(I, n)
        jlo, jhi = _bounds(J, n)
        for i in range(ilo, ihi + 1):
            for j in range(max(jlo, i), jhi + 1):
                out[(i, j)] = val
    return out


def soft_cutoffs() -> np.ndarray:
    """Symmetric MAXT x MAXT matrix of ``pair_style soft`` cutoffs.

    run_single_shoebox.py:589-625 == VisitorGeneMaps/run_single_GENES_STAGES.py:569-599.
    """
    sig_chr, sig_rbp, sig_rnp, sig_ser5p = 1.0, 0.3, 1.0, 0.3
    w = 2 ** (1 / 6)
    c_cc = w
    c_crb = ((sig_chr + sig_rbp) / 2) * w
    c_crn = ((sig_chr + sig_rnp) / 2) * w
    c_cs = ((sig_chr + sig_ser5p) / 2) * w
    c_rbrn = ((sig_rbp + sig_rnp) / 2) * w
    c_rbs = ((sig_rbp + sig_ser5p) / 2) * w
    c_rns = ((sig_rnp + sig_ser5p) / 2) * w
    stream = [
        ("*", "*", w),
        ("1*2", "4", c_crb), ("1*2", "5", 1.5 * c_crn), ("1*2", "6", c_cc), ("1*2", "7", 1.5 * c_cc), ("1*2", "8", 1.0 * c_cs),
        ("4", "4", sig_rbp * c_cc), ("4", "5", c_rbrn), ("4", "6*7", c_crb), ("4", "8", c_rbs), ("4", "9*11", c_crb),
        ("5", "5", sig_rnp * c_cc), ("5", "6", 1.5 * c_crn), ("5", "7", c_crn), ("5", "8", c_rns), ("5", "9*11", 1.5 * c_crn),
        ("6", "7", 1.5 * c_cc), ("6", "8", c_cs), ("6", "9*11", c_cc),
        ("7", "7", 1.5 * c_cc), ("7", "8", 1.0 * c_cs), ("7", "9*11", 1.5 * c_cc),
        ("8", "8", sig_ser5p * c_cc), ("8", "9*11", c_cs),
    ]
    m = np.zeros((MAXT, MAXT))
    for (i, j), v in _replay(stream).items():
        m[i, j] = m[j, i] = v
    return m


# --- pair_style table streams: (I, J, (KEY, cut)) -----------------------------------------------
def _table_stream_phase_a(driver: str, condition: str):
    """Commands issued when (i+1) == t_soft.

    shoebox: run_single_shoebox.py:803-855; genes_stages: VGM/run_single_GENES_STAGES.py:745-779.
    """
    s = [("*", "*", ("CHROMATIN_CHROMATIN", 1.2)),
         ("1*2", "3", ("ACTIN_CHROMATIN", 1.2)),
         ("1*2", "4", ("RBP_CHROMATIN", 0.75)),
         ("1*2", "5", ("RNP_CHROMATIN", 1.75)),
         ("1*2", "6", ("INDUCED_CHROMATIN", 1.2)),
         ("1*2", "7", ("ACTIVE_CHROMATIN", 1.75))]
    if driver == "shoebox":
        if condition in ("SwinA_old", "SwinA_nopol"):
            s.append(("3", "3", ("ACTIN_ACTIN_STRONGATTRACTION", 0.9)))
        else:
            s.append(("3", "3", ("ACTIN_ACTIN", 0.4)))
        s += [("3", "4", ("ACTIN_RBP", 0.4)), ("3", "5", ("RNP_ACTIN", 1.9)),
              ("3", "6", ("ACTIN_INDUCED", 1.2)), ("3", "7", ("ACTIN_ACTIVE", 1.2)),
              ("3", "9", ("ACTIN_REGULATORY", 0.75)), ("3", "10*11", ("ACTIN_PROMOTER", 0.75))]
    s += [("4", "4", ("RBP_RBP", 0.4)), ("4", "5", ("RNP_RBP", 0.75)), ("4", "6", ("RBP_INDUCED", 0.75)),
          ("4", "7", ("RBP_ACTIVE", 0.75)), ("4", "8", ("RBP_SER5P", 0.4)), ("4", "9*11", ("RBP_REGULATORY", 0.75)),
          ("5", "5", ("RNP_RNP_ATTRACTION", 2.9)), ("5", "6", ("RNP_INDUCED", 1.75)), ("5", "7", ("RNP_ACTIVE", 2.9)),
          ("5", "8", ("RNP_SER5P", 0.75)), ("5", "9", ("RNP_REGULATORY", 1.75)), ("5", "10*11", ("RNP_PROMOTER", 1.75)),
          ("6", "7", ("INDUCED_ACTIVE", 1.75)),
          ("7", "7", ("ACTIVE_ACTIVE", 1.75)), ("7", "9", ("ACTIVE_REGULATORY", 1.75)),
          # Ser5P behaves like RBP until it is switched on
          ("1*2", "8", ("RBP_CHROMATIN", 0.75)), ("3", "8", ("ACTIN_RBP", 0.4)), ("6", "8", ("RBP_INDUCED", 0.75)),
          ("7", "8", ("RBP_ACTIVE", 0.75)), ("8", "8", ("RBP_SER5P", 0.4)), ("8", "9*11", ("RBP_REGULATORY", 0.75))]
    return s


def _table_stream_phase_b(driver: str, condition: str):
    """Extra commands when (i+1) == t_ser5p_on.

    shoebox: run_single_shoebox.py:862-874; genes_stages: VGM/run_single_GENES_STAGES.py:785-791.
    """
    s = [("8", "8", ("SER5P_SER5P", 0.9))]
    if driver == "shoebox":
        if condition == "noActinSer5P":
            s.append(("3", "8", ("ACTIN_SER5P_NOATTRACTION", 0.4)))
        else:
            s.append(("3", "8", ("ACTIN_SER5P", 0.9)))
    s += [("6", "8", ("INDUCED_SER5P", 0.75)), ("7", "8", ("ACTIVE_SER5P", 1.2)), ("1*2", "8", ("CHROMATIN_SER5P", 1.5)),
          ("8", "9", ("SER5P_REGULATORY_WEAK", 1.9)), ("8", "10", ("PROMOTER_SER5P_STRONG", 1.9)),
          ("8", "11", ("ACTIVEPROMOTER_SER5P", 1.2))]
    return s


def _table_stream_treatment(condition: str):
    """Commands at i == NRuns (run_single_shoebox.py:876-887)."""
    if condition == "Hexanediol":
        return [("8", "8", ("RBP_RBP", 0.4)), ("8", "9", ("SER5P_REGULATORY_JQ1", 2.0))]
    if condition == "JQ-1":
        return [("8", "9", ("RBP_REGULATORY", 0.75))]
    return []


def table_map(driver: str = "genes_stages", condition: str = "Control", phase: str = "B") -> dict:
    """Effective {(i,j): (KEY, cut)} for i <= j after phase "A", "B" or "T" (treatment)."""
    if driver not in ("genes_stages", "shoebox"):
        raise ValueError(f"unknown driver {driver!r}")
    s = _table_stream_phase_a(driver, condition)
    if phase in ("B", "T"):
        s = s + _table_stream_phase_b(driver, condition)
    if phase == "T":
        s = s + _table_stream_treatment(condition)
    return _replay(s)


@dataclass
class PairTables:
    """Distinct (KEY, cut) tables + the per-phase type-pair -> table-index maps."""

    specs: list            # [(KEY, cut)]
    maps: dict             # phase -> int32 MAXT x MAXT (symmetric; -1 where unused)


def pair_tables(driver: str = "genes_stages", condition: str = "Control", phases=("A", "B")) -> PairTables:
    specs: list = []
    maps = {}
    for ph in phases:
        m = np.full((MAXT, MAXT), -1, np.int32)
        for (i, j), spec in table_map(driver, condition, ph).items():
            if spec not in specs:
                specs.append(spec)
            m[i, j] = m[j, i] = specs.index(spec)
        maps[ph] = m
    return PairTables(specs, maps)


@dataclass(frozen=True)
class ChunkPlan:
    """What one Python timestep i does before and during its ``run``."""

    index: int                  # i (0-based)
    pair: str                   # "soft" | "A" | "B" | "T"
    ramp: tuple | None          # (start, stop) of `run 200 start S stop E`, or None
    adapt_soft: bool            # fix s1 alive
    adapt_bond1: bool           # fix s2 alive
    adapt_bond2: bool           # fix s3 alive (shoebox driver only; never unfixed, :638-639 vs :794-795)
    n_steps: int = T_RUN


def chunk_schedule(driver: str = "genes_stages", condition: str = "Control", n_chunks: int | None = None,
                   is_actin_simulation: bool = False) -> list:
    """Per-chunk plan for i in range(NRuns + treatment_duration).

    run_single_shoebox.py:759, 793-899; VGM/run_single_GENES_STAGES.py:708, 738-820.
    """
    total = N_RUNS + TREATMENT_DURATION.get(condition, 0)
    if n_chunks is None:
        n_chunks = total
    plans = []
    for i in range(n_chunks):
        soft = (i + 1) < T_SOFT
        if soft:
            pair = "soft"
        elif (i + 1) < T_SER5P_ON:
            pair = "A"
        elif i >= N_RUNS and condition in ("Hexanediol", "JQ-1"):
            pair = "T"
        else:
            pair = "B"
        plans.append(ChunkPlan(
            index=i, pair=pair, ramp=(0, T_SOFT * T_RUN) if soft else None,
            adapt_soft=soft, adapt_bond1=soft,
            adapt_bond2=(driver == "shoebox"),
        ))
    return plans
