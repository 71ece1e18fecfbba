"""The LJsoft tabulated pair potential: closed form, section parameters, table-file I/O.

Restates ``ljsoft_potential.py`` (== ``setup_ljsoft_potentials.py``) of the reference:

* closed form ``compute_force`` / ``compute_energy`` (ljsoft_potential.py:10-27)
* section geometry ``write_section`` (ljsoft_potential.py:81-108): eps = eps0/lam^n,
  s = a*s0*gam^(1/6), rc = b*s0*delt^(1/6), delt = 2 - alpha(1-lam)^2, gam = 2/delt; r grid
  ``sqrt(linspace(dr^2, rmax^2, N))``; header ``N <N> RSQ <r0> <rN>``
* base parameters lam=0.2, n=1, alpha=0.5, N=30000, dr=1e-4, rmax=3.0 (:143-150)
* the 55 keyword sections and their (eps0, s0, a, b) (:181-340)

The 57 MB ``LJsoft_RSQ.table`` the drivers read from the working directory
(run_single_shoebox.py:803) is not in the reference checkout (``.MISSING_LARGE_BLOBS``);
`write_table_file` regenerates it, `read_table_file` reads any file of that dialect.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LAM, NEXP, ALPHA = 0.2, 1, 0.5
N_ROWS, DR, RMAX = 30000, 0.0001, 3.0

_sig = dict(actin=0.3, chr=1.0, rbp=0.3, rnp=1.0, ser5p=0.3)
S_CA = (_sig["chr"] + _sig["actin"]) / 2
S_CRB = (_sig["chr"] + _sig["rbp"]) / 2
S_CRN = (_sig["chr"] + _sig["rnp"]) / 2
S_CS = (_sig["chr"] + _sig["ser5p"]) / 2
S_ARB = (_sig["actin"] + _sig["rbp"]) / 2
S_ARN = (_sig["actin"] + _sig["rnp"]) / 2
S_AS = (_sig["actin"] + _sig["ser5p"]) / 2
S_RBRN = (_sig["rbp"] + _sig["rnp"]) / 2
S_RBS = (_sig["rbp"] + _sig["ser5p"]) / 2
S_RNS = (_sig["rnp"] + _sig["ser5p"]) / 2
_gRE, _fRE = 1, 1.5

# (KEY, eps0, s0, a, b, longrange) in file order -- ljsoft_potential.py:181-340
SECTIONS: list[tuple[str, float, float, float, float, bool]] = [
    ("CHROMATIN_CHROMATIN", _gRE, 1, 1, 1, False),
    ("CHROMATIN_SER5P", _fRE, S_CS, 2.0, 2.0, False),
    ("CHROMATIN_REGULATORY", _gRE, 1, 1, 1, False),
    ("ACTIN_CHROMATIN", _fRE, S_CA, 1.5, 1.5, False),
    ("ACTIN_ACTIN_WEAKATTRACTION", 1.0 / 2, 0.3, 1.0, 2.5, False),
    ("ACTIN_ACTIN_ATTRACTION", 1.0, 0.3, 1.0, 2.5, False),
    ("ACTIN_ACTIN_STRONGATTRACTION", 1.5 * 1.0, 0.3, 1.0, 2.5, False),
    ("ACTIN_ACTIN", _gRE, 0.3, 1, 1, False),
    ("ACTIN_RBP", _gRE, S_ARB, 1, 1, False),
    ("ACTIN_INDUCED", _fRE, S_CA, 1.5, 1.5, False),
    ("ACTIN_ACTIVE", _fRE, S_CA, 1.5, 1.5, False),
    ("ACTIN_SER5P", 1.5, S_AS, 1, 2.5, False),
    ("ACTIN_SER5P_LONGRANGE", 1, S_AS, 1, 2.5, True),
    ("ACTIN_SER5P_WEAKATTRACTION", 1.2, S_AS, 1, 2.5, False),
    ("ACTIN_SER5P_VERYWEAKATTRACTION", 1.0, S_AS, 1, 2.5, False),
    ("ACTIN_SER5P_NOATTRACTION", _gRE, S_AS, 1, 1, False),
    ("ACTIN_REGULATORY", _gRE, S_CA, 1.0, 1.0, False),
    ("ACTIN_PROMOTER", _gRE, S_CA, 1.0, 1.0, False),
    ("INDUCED_CHROMATIN", _fRE, 1, 1.5, 1.5, False),
    ("INDUCED_INDUCED", _gRE, 1, 1, 1, False),
    ("INDUCED_ACTIVE", _fRE, 1, 1.5, 1.5, False),
    ("INDUCED_SER5P", _gRE, S_CS, 1, 1.0, False),
    ("INDUCED_REGULATORY", _gRE, 1, 1, 1, False),
    ("ACTIVE_CHROMATIN", _fRE, 1, 1.5, 1.5, False),
    ("ACTIVE_ACTIVE", _fRE, 1, 1.5, 1.5, False),
    ("ACTIVE_SER5P", _fRE, S_CS, 1.5, 1.5, False),
    ("ACTIVE_REGULATORY", _fRE, 1, 1.5, 1.5, False),
    ("RBP_CHROMATIN", _gRE, S_CRB, 1, 1, False),
    ("RBP_RBP", _gRE, 0.3, 1, 1, False),
    ("RBP_INDUCED", _gRE, S_CRB, 1, 1, False),
    ("RBP_ACTIVE", _gRE, S_CRB, 1, 1, False),
    ("RBP_SER5P", _gRE, S_RBS, 1, 1, False),
    ("RBP_REGULATORY", _gRE, S_CRB, 1, 1, False),
    ("RNP_CHROMATIN", _fRE, S_CRN, 1.5, 1.5, False),
    ("RNP_ACTIN", _gRE, S_ARN, 1, 1, False),
    ("RNP_ACTIN_ATTR", 1, S_ARN, 1, 1.5, False),
    ("RNP_RBP", _gRE, S_RBRN, 1, 1, False),
    ("RNP_RNP_HCREPULSIVE", _gRE, 1.0, 1, 1, False),
    ("RNP_RNP_ATTRACTION", 1, 1.0, 1, 2.5, False),
    ("RNP_INDUCED", _fRE, S_CRN, 1.5, 1.5, False),
    ("RNP_ACTIVE", 1, S_CRN, 1, 2.5, False),
    ("RNP_SER5P", _gRE, S_RNS, 1, 1, False),
    ("RNP_REGULATORY", _fRE, S_CRN, 1.5, 1.5, False),
    ("RNP_PROMOTER", _fRE, S_CRN, 1.5, 1.5, False),
    ("SER5P_SER5P", 1.0, 0.3, 1, 2.5, False),
    ("SER5P_REGULATORY", 2.0, S_CS, 1, 2.5, False),
    ("SER5P_REGULATORY_STRONG", 1.75, S_CS, 1, 2.5, False),
    ("SER5P_REGULATORY_MEDIUM", 1.5, S_CS, 1, 2.5, False),
    ("SER5P_REGULATORY_WEAK", 1.25, S_CS, 1, 2.5, False),
    ("SER5P_REGULATORY_VERYWEAK", 1, S_CS, 1, 2.5, False),
    ("SER5P_REGULATORY_JQ1", 2.0 / 2, S_CS, 1, 2.5, False),
    ("REGULATORY_REGULATORY", _gRE, 1, 1, 1, False),
    ("PROMOTER_SER5P", 1.5, S_CS, 1, 2.5, False),
    ("PROMOTER_SER5P_STRONG", 2.0, S_CS, 1, 2.5, False),
    ("ACTIVEPROMOTER_SER5P", _fRE, S_CS, 1.5, 1.5, False),
]
SECTION_INDEX = {s[0]: k for k, s in enumerate(SECTIONS)}


@dataclass(frozen=True)
class SectionGeometry:
    """Derived closed-form constants of one section (ljsoft_potential.py:82-86)."""

    key: str
    eps: float      # eps0 / lam^n
    eps_eff: float  # lam^n * eps = eps0: the prefactor that actually multiplies the force
    s: float
    rc: float
    c: float        # alpha (1-lam)^2
    longrange: bool


def section_geometry(key: str) -> SectionGeometry:
    _, eps0, s0, a, b, lr = SECTIONS[SECTION_INDEX[key]]
    delt = 2 - ALPHA * (1 - LAM) ** 2
    gam = 2 / delt
    eps = eps0 / LAM ** NEXP
    s = a * s0 * gam ** (1 / 6)
    rc = b * s0 * delt ** (1 / 6)
    return SectionGeometry(key, eps, (LAM ** NEXP) * eps, s, rc, ALPHA * (1 - LAM) ** 2, lr)


def r_grid(n: int = N_ROWS, dr: float = DR, rmax: float = RMAX) -> np.ndarray:
    """ljsoft_potential.py:93."""
    return np.sqrt(np.linspace(dr ** 2, rmax ** 2, n))


def _force_energy_scalar(l, n, eps, a, s, rc, r):
    """Scalar closed form, operation for operation as ljsoft_potential.py:10-27 (Python floats)."""
    xc = a * (1 - l) ** 2 + (rc / s) ** 6
    Ec = (l ** n) * 4 * eps * (1 / (xc ** 2) - 1 / xc)
    if r < rc:
        x = a * (1 - l) ** 2 + (r / s) ** 6
        F = 24 * (l ** n) * eps * (r ** 5) / (s ** 6) * (2 / (x ** 3) - 1 / (x ** 2))
        E = (l ** n) * 4 * eps * (1 / (x ** 2) - 1 / x)
        E = E - Ec
    else:
        F = 0
        E = 0
    return E, F


def _force_energy_longrange_scalar(l, n, eps, a, s, rc, r):
    """ljsoft_potential.py:29-62 (linear tail between the minimum and rc)."""
    xc = a * (1 - l) ** 2 + (rc / s) ** 6
    Ec = (l ** n) * 4 * eps * (1 / (xc ** 2) - 1 / xc)
    rmin = s * (2 - a * (1 - l) ** 2) ** (1 / 6)
    xmin = a * (1 - l) ** 2 + (rmin / s) ** 6
    Emin = (l ** n) * 4 * eps * (1 / (xmin ** 2) - 1 / xmin)
    m = (Emin - Ec) / (rmin - rc)
    c = -m * rc
    if r < rmin:
        x = a * (1 - l) ** 2 + (r / s) ** 6
        F = 24 * (l ** n) * eps * (r ** 5) / (s ** 6) * (2 / (x ** 3) - 1 / (x ** 2))
        E = (l ** n) * 4 * eps * (1 / (x ** 2) - 1 / x) - Ec
    elif r < rc:
        F = -m
        E = m * r + c
    else:
        F = 0
        E = 0
    return E, F


def _section_rows(key: str, n: int = N_ROWS, dr: float = DR, rmax: float = RMAX):
    """Yield (r, E, F) per row as the Python scalars the reference's writer sees.

    Beyond rc the reference assigns the *int* 0 (ljsoft_potential.py:15,26), which prints as
    "0"; inside rc the values are floats.
    """
    g = section_geometry(key)
    for rk in r_grid(n, dr, rmax):
        rk = float(rk)
        e, f = _force_energy_scalar(LAM, NEXP, g.eps, ALPHA, g.s, g.rc, rk)
        if g.longrange:
            e2, f2 = _force_energy_longrange_scalar(LAM, NEXP, g.eps, ALPHA, g.s, g.rc, rk)
            e, f = (e2 + e) / 2, (f2 + f) / 2
        yield rk, e, f


def section_columns(key: str, n: int = N_ROWS, dr: float = DR, rmax: float = RMAX):
    """(r, E, F) float64 columns of one section with the values the reference writes."""
    rows = list(_section_rows(key, n, dr, rmax))
    return (np.array([r[0] for r in rows]), np.array([float(r[1]) for r in rows]),
            np.array([float(r[2]) for r in rows]))


def write_table_file(path: str, keys: list[str] | None = None) -> None:
    """Regenerate ``LJsoft_RSQ.table`` (all 55 sections by default) in the reference's layout
    (ljsoft_potential.py:95-108 and the blank line between sections, :182)."""
    keys = [s[0] for s in SECTIONS] if keys is None else keys
    with open(path, "w") as f:
        for q, key in enumerate(keys):
            rows = list(_section_rows(key))
            f.write(key + "\n")
            f.write("N\t" + str(len(rows)) + "\t" + "RSQ" + "\t" + str(rows[0][0]) + "\t" + str(rows[-1][0]) + "\n")
            f.write("\n")
            for k, (r, e, fo) in enumerate(rows):
                f.write(str(k + 1) + "\t" + str(r) + "\t" + str(e) + "\t" + str(fo) + "\n")
            if q != len(keys) - 1:
                f.write("\n")


@dataclass
class TableSection:
    key: str
    n: int
    rflag: str            # "RSQ" (the only style the reference writes)
    rlo: float
    rhi: float
    r: np.ndarray
    e: np.ndarray
    f: np.ndarray


def read_table_file(path: str, keys: set[str] | None = None) -> dict[str, TableSection]:
    """Read sections of a LAMMPS pair-table file ([LAMMPS-ext] TableFileReader semantics).

    A section is: keyword line, parameter line ``N n [R|RSQ lo hi]``, blank line, n rows
    ``index r E F``.  Only `keys` are parsed when given (the file is large).
    """
    out: dict[str, TableSection] = {}
    with open(path, "r") as fh:
        it = iter(fh)
        for line in it:
            word = line.strip()
            if not word or word.startswith("#"):
                continue
            key = word.split()[0]
            params = next(it).split()
            if params[0] != "N":
                raise ValueError(f"{path}: section {key}: expected 'N', got {params[:1]}")
            n = int(params[1])
            rflag, rlo, rhi = "", 0.0, 0.0
            if len(params) >= 5 and params[2] in ("R", "RSQ"):
                rflag, rlo, rhi = params[2], float(params[3]), float(params[4])
            next(it)  # blank line
            if keys is not None and key not in keys:
                for _ in range(n):
                    next(it)
                continue
            arr = np.loadtxt([next(it) for _ in range(n)], dtype=np.float64, ndmin=2)
            out[key] = TableSection(key, n, rflag, rlo, rhi, arr[:, 1].copy(), arr[:, 2].copy(), arr[:, 3].copy())
    return out


def generated_section(key: str) -> TableSection:
    """The section `key` as `write_table_file` would write it, without touching the disk."""
    r, E, F = section_columns(key)
    # the file stores str(float); parsing it back is the identity for Python floats
    return TableSection(key, len(r), "RSQ", float(str(r[0])), float(str(r[-1])), r, E, F)
