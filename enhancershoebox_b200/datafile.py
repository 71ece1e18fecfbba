"""LAMMPS data-file dialect of the reference's input builder, and a seeded IC generator.

Reads (and writes) exactly the `atom_style angle` data files that the reference's
``make_input_file.py:186-328`` produces and that ``read_data`` consumes at
``run_single_shoebox.py:531`` / ``VisitorGeneMaps/run_single_GENES_STAGES.py:537``:
header counts, box, ``Masses``, ``Atoms`` (``id mol type x y z ix iy iz``), ``Velocities``,
``Bonds`` (``id type a b``), ``Angles`` (``id type a b c``).

`make_shoebox` builds the same initial condition from a seed (the reference never seeds its
RNG, ``make_input_file.py:11-15,37,248-269``); it exists so tests and benches can make
reproducible replicas on a machine that has no reference checkout.  It is not a copy of the
reference's generator: same geometry and counts, numpy ``Generator`` streams, no file I/O.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# atom types (run_single_shoebox.py:23-33)
CHROMATIN, INACTIVE_GENE, ACTIN, RBP, RNP, INDUCED_GENE, ACTIVE_GENE, SER5P, ENHANCER, PROMOTER, ACTIVE_PROMOTER = range(1, 12)
N_ATOM_TYPES = 11

# per-box species counts (make_input_file.py:73-117; run_single_shoebox.py:334-369)
BOX_TABLE = {
    9: dict(polymer=300, actin=240, ser5p=300, rbp=450),
    10: dict(polymer=336, actin=270, ser5p=336, rbp=500),
    11: dict(polymer=370, actin=300, ser5p=370, rbp=550),
    12: dict(polymer=400, actin=300, ser5p=400, rbp=600),
    13: dict(polymer=433, actin=325, ser5p=433, rbp=650),
    15: dict(polymer=500, actin=400, ser5p=500, rbp=750),
    18: dict(polymer=600, actin=500, ser5p=600, rbp=900),
    21: dict(polymer=700, actin=560, ser5p=700, rbp=1050),
    24: dict(polymer=800, actin=650, ser5p=800, rbp=1200),
}
LENGTH_ENHANCER = 50  # make_input_file.py:125
LENGTH_GENE = 5       # make_input_file.py:129
BOX_Y, BOX_Z, BOX_PAD = 7.5, 5.0, 2.0  # make_input_file.py:182-184


@dataclass
class ShoeboxData:
    """One replica's system as read from a data file.  Atom order is by atom ID (1..N)."""

    types: np.ndarray                 # int32 [N], 1-based atom types
    x: np.ndarray                     # float64 [N,3]
    v: np.ndarray                     # float64 [N,3]
    mol: np.ndarray                   # int32 [N]
    bonds: np.ndarray                 # int32 [nb,3]  (bond type, i, j), 0-based atom indices
    angles: np.ndarray                # int32 [na,4]  (angle type, i, j, k), j the vertex
    box_lo: np.ndarray                # float64 [3]
    box_hi: np.ndarray                # float64 [3]
    n_atom_types: int = N_ATOM_TYPES
    masses: np.ndarray = field(default_factory=lambda: np.ones(N_ATOM_TYPES))
    title: str = ""

    @property
    def n_atoms(self) -> int:
        return int(self.types.shape[0])

    # ---- derived quantities the drivers compute right after read_data --------------------
    def anchors(self) -> np.ndarray:
        """0-based indices of the frozen polymer ends.

        The reference hard-codes them per polymer length (``run_single_shoebox.py:651-668``:
        {1, N/3, N/3+1, N} as atom IDs); that is the first and last bead of each of the two
        chromatin polymers, which is how they are derived here (molecules 1 and 2).
        """
        out = []
        for m in (1, 2):
            idx = np.nonzero(self.mol == m)[0]
            idx = idx[np.isin(self.types[idx], [ACTIN, RBP, RNP, SER5P], invert=True)]
            if idx.size:
                out += [int(idx[0]), int(idx[-1])]
        return np.asarray(out, dtype=np.int32)

    def enhancer_atoms(self) -> np.ndarray:
        return np.nonzero(self.types == ENHANCER)[0].astype(np.int32)

    def ser5p_atoms(self) -> np.ndarray:
        return np.nonzero(self.types == SER5P)[0].astype(np.int32)

    def rbp_atoms(self) -> np.ndarray:
        return np.nonzero(np.isin(self.types, [RBP, RNP]))[0].astype(np.int32)

    def gene_start(self) -> int:
        """0-based index of the first bead of the (single) gene body.

        ``run_single_shoebox.py:683-694``: first atom of type 2 whose predecessor is not type 2.
        Only one gene is ever generated (``make_input_file.py:128``).
        """
        g = np.nonzero(self.types == INACTIVE_GENE)[0]
        if g.size == 0:
            raise ValueError("data file has no inactive-gene beads (type 2)")
        starts = [int(a) for a in g if a == 0 or self.types[a - 1] != INACTIVE_GENE]
        if len(starts) != 1:
            raise ValueError(f"expected exactly one gene, found {len(starts)}")
        return starts[0]

    def promoter_length(self) -> int:
        g = self.gene_start()
        p = 0
        while g - 1 - p >= 0 and self.types[g - 1 - p] == PROMOTER:
            p += 1
        return p


def read_data(path: str) -> ShoeboxData:
    """Parse a data file written by ``make_input_file.py:186-328`` (any title line)."""
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    title = lines[0]
    counts = {}
    box = {}
    sections = {}
    i = 1
    header_keys = ("atoms", "bonds", "angles", "atom types", "bond types", "angle types")
    section_names = ("Masses", "Atoms", "Velocities", "Bonds", "Angles")
    cur = None
    while i < len(lines):
        ln = lines[i].split("#")[0].strip()
        i += 1
        if not ln:
            continue
        first = ln.split()[0]
        if first in section_names:
            cur = first
            sections[cur] = []
            continue
        if cur is None:
            matched = False
            for key in sorted(header_keys, key=len, reverse=True):
                if ln.endswith(" " + key):
                    counts[key] = int(ln.split()[0])
                    matched = True
                    break
            if matched:
                continue
            for ax in ("x", "y", "z"):
                if ln.endswith(f"{ax}lo {ax}hi"):
                    a, b = ln.split()[:2]
                    box[ax] = (float(a), float(b))
                    matched = True
            if not matched:
                raise ValueError(f"{path}: unrecognised header line: {ln!r}")
        else:
            sections[cur].append(ln.split())
    n = counts["atoms"]
    atoms = sections.get("Atoms", [])
    if len(atoms) != n:
        raise ValueError(f"{path}: header says {n} atoms, Atoms section has {len(atoms)}")
    types = np.zeros(n, np.int32)
    mol = np.zeros(n, np.int32)
    x = np.zeros((n, 3))
    v = np.zeros((n, 3))
    for row in atoms:
        k = int(row[0]) - 1
        mol[k] = int(row[1])
        types[k] = int(row[2])
        x[k] = [float(row[3]), float(row[4]), float(row[5])]
    for row in sections.get("Velocities", []):
        v[int(row[0]) - 1] = [float(row[1]), float(row[2]), float(row[3])]
    nb, na = counts.get("bonds", 0), counts.get("angles", 0)
    bonds = np.array([[int(r[1]), int(r[2]) - 1, int(r[3]) - 1] for r in sections.get("Bonds", [])], np.int32).reshape(-1, 3)
    angles = np.array([[int(r[1]), int(r[2]) - 1, int(r[3]) - 1, int(r[4]) - 1] for r in sections.get("Angles", [])], np.int32).reshape(-1, 4)
    if bonds.shape[0] != nb or angles.shape[0] != na:
        raise ValueError(f"{path}: bond/angle counts do not match the header")
    nt = counts.get("atom types", N_ATOM_TYPES)
    masses = np.ones(nt)
    for row in sections.get("Masses", []):
        masses[int(row[0]) - 1] = float(row[1])
    return ShoeboxData(
        types=types, x=x, v=v, mol=mol, bonds=bonds, angles=angles,
        box_lo=np.array([box["x"][0], box["y"][0], box["z"][0]]),
        box_hi=np.array([box["x"][1], box["y"][1], box["z"][1]]),
        n_atom_types=nt, masses=masses, title=title,
    )


def write_data(path: str, d: ShoeboxData) -> None:
    """Write `d` in the dialect ``make_input_file.py:186-328`` uses (readable by LAMMPS read_data)."""
    n = d.n_atoms
    with open(path, "w") as f:
        f.write((d.title or "LAMMPS input file for a SHOEBOX simulation (shoebox-b200 generator).") + "\n\n")
        f.write(f"{n} atoms\n{d.bonds.shape[0]} bonds\n{d.angles.shape[0]} angles\n\n")
        f.write(f"{d.n_atom_types} atom types\n2 bond types\n2 angle types\n\n")
        for ax, k in zip("xyz", range(3)):
            f.write(f"{d.box_lo[k]!r} {d.box_hi[k]!r} {ax}lo {ax}hi\n".replace("np.float64(", "").replace(")", ""))
        f.write("\nMasses\n\n")
        for t in range(d.n_atom_types):
            f.write(f"{t + 1} {d.masses[t]:g}\n")
        f.write("\nAtoms\n\n")
        for k in range(n):
            f.write(f"{k + 1} {int(d.mol[k])} {int(d.types[k])} {float(d.x[k, 0])!r} {float(d.x[k, 1])!r} {float(d.x[k, 2])!r} 0 0 0\n")
        f.write("\nVelocities\n\n")
        for k in range(n):
            f.write(f"{k + 1} {float(d.v[k, 0])!r} {float(d.v[k, 1])!r} {float(d.v[k, 2])!r}\n")
        f.write("\nBonds\n\n")
        for k, (bt, a, b) in enumerate(d.bonds):
            f.write(f"{k + 1} {bt} {a + 1} {b + 1}\n")
        f.write("\nAngles\n\n")
        for k, (at, a, b, c) in enumerate(d.angles):
            f.write(f"{k + 1} {at} {a + 1} {b + 1} {c + 1}\n")


def _polymer_types(box_size: int, promoter_length: int) -> tuple[np.ndarray, int]:
    """Type vector of the two chromatin polymers (make_input_file.py:123-154).

    Left polymer (1/3 of the beads): a 50-bead enhancer centred in it.  Right polymer: one 5-bead
    gene centred in it, preceded by `promoter_length` promoter beads.  Single elements are always
    placed at the middle of the carrier (``put_stuff_on``, make_input_file.py:36-39).
    """
    npoly = BOX_TABLE[box_size]["polymer"]
    n_left = npoly // 3
    n_right_plain = npoly - n_left - promoter_length

    def place(total, elem_type, elem_len):
        sub = total - elem_len
        pos = sub // 2
        return [CHROMATIN] * (pos + 1) + [elem_type] * elem_len + [CHROMATIN] * (sub - pos - 1)

    left = place(n_left, ENHANCER, LENGTH_ENHANCER)
    right = place(n_right_plain, INACTIVE_GENE, LENGTH_GENE)
    g = right.index(INACTIVE_GENE)
    right = right[:g] + [PROMOTER] * promoter_length + right[g:]
    assert len(left) + len(right) == npoly
    return np.array(left + right, np.int32), n_left


def make_shoebox(box_size: int = 11, promoter_length: int = 3, is_actin_simulation: bool = False,
                 seed: int = 0, filament_length: int = 3, number_of_filaments: int | None = None) -> ShoeboxData:
    """Seeded equivalent of ``make_input_file_general`` (make_input_file.py:51-329).

    Polymers start as straight lines on the +-box x-faces along y in [-3.75, 3.75], gene and
    promoter beads at x = 0 (:214-234); RBP uniform in the inner box (:254-258); actin monomers
    and Ser5P uniform in the left half (:247-251, :268-272); free species rounded to 6 decimals;
    zero velocities (:277-280); type-1 bonds/angles along each polymer (:289-319), a type-2
    bonded+angled seed filament when actin is present (:239-244, :299-328).
    """
    cfg = BOX_TABLE[box_size]
    rng = np.random.default_rng(np.random.SeedSequence([0x5B0E, box_size, promoter_length, int(is_actin_simulation), seed]))
    npoly = cfg["polymer"]
    n_actin = cfg["actin"] if is_actin_simulation else 0
    if number_of_filaments is None:
        number_of_filaments = 1 if is_actin_simulation else 0
    if not is_actin_simulation:
        number_of_filaments = 0
    n_fil = number_of_filaments * filament_length
    n = npoly + n_fil + n_actin + cfg["rbp"] + cfg["ser5p"]
    poly_types, n_left = _polymer_types(box_size, promoter_length)

    types = np.zeros(n, np.int32)
    mol = np.zeros(n, np.int32)
    x = np.zeros((n, 3))
    types[:npoly] = poly_types
    bx = float(box_size)
    x[:n_left, 0] = -bx
    x[:n_left, 1] = np.linspace(-BOX_Y / 2, BOX_Y / 2, n_left)
    mol[:n_left] = 1
    n_right = npoly - n_left
    x[n_left:npoly, 1] = np.linspace(-BOX_Y / 2, BOX_Y / 2, n_right)
    right_t = poly_types[n_left:]
    x[n_left:npoly, 0] = np.where(np.isin(right_t, [INACTIVE_GENE, PROMOTER]), 0.0, bx)
    mol[n_left:npoly] = 2
    k = npoly
    m = 3
    for _ in range(number_of_filaments):
        # make_input_file.py:21-29: every monomer is drawn around the origin in a shell [0.2, 0.4]
        for _j in range(filament_length):
            while True:
                p = rng.uniform(-0.4, 0.4, 3)
                dist = float(np.sqrt(np.sum(p * p)))
                if 0.2 <= dist <= 0.4:
                    break
            x[k] = np.round(p, 6)
            types[k] = ACTIN
            mol[k] = m
            k += 1
        m += 1

    def scatter(count, lo, hi, t):
        nonlocal k, m
        pts = np.round(rng.uniform(lo, hi, size=(count, 3)), 6)
        x[k:k + count] = pts
        types[k:k + count] = t
        mol[k:k + count] = np.arange(m, m + count)
        k += count
        m += count

    scatter(n_actin, [-bx, -BOX_Y, -BOX_Z], [0, BOX_Y, BOX_Z], ACTIN)
    scatter(cfg["rbp"], [-bx, -BOX_Y, -BOX_Z], [bx, BOX_Y, BOX_Z], RBP)
    scatter(cfg["ser5p"], [-bx, -BOX_Y, -BOX_Z], [0, BOX_Y, BOX_Z], SER5P)
    assert k == n

    bonds, angles = [], []
    for a in range(npoly - 1):
        if a == n_left - 1:
            continue
        bonds.append((1, a, a + 1))
    for a in range(npoly - 2):
        if a in (n_left - 1, n_left - 2):
            continue
        angles.append((1, a, a + 1, a + 2))
    a0 = npoly
    for _ in range(number_of_filaments):
        for j in range(filament_length - 1):
            bonds.append((2, a0 + j, a0 + j + 1))
        for j in range(filament_length - 2):
            angles.append((2, a0 + j, a0 + j + 1, a0 + j + 2))
        a0 += filament_length
    return ShoeboxData(
        types=types, x=x, v=np.zeros((n, 3)), mol=mol,
        bonds=np.array(bonds, np.int32).reshape(-1, 3), angles=np.array(angles, np.int32).reshape(-1, 4),
        box_lo=np.array([-bx - BOX_PAD, -BOX_Y - BOX_PAD, -BOX_Z - BOX_PAD]),
        box_hi=np.array([bx + BOX_PAD, BOX_Y + BOX_PAD, BOX_Z + BOX_PAD]),
        title=f"shoebox-b200 IC box={box_size} promoter={promoter_length} actin={int(is_actin_simulation)} seed={seed}",
    )


def input_file_name(box_size: int, promoter_length: int, is_actin_simulation: bool, init_free_actin: int | None = None) -> str:
    """File name the drivers look for (run_single_shoebox.py:449-453; make_input_file.py:158-161)."""
    cfg = BOX_TABLE[box_size]
    if is_actin_simulation:
        z = cfg["actin"] if init_free_actin is None else init_free_actin
        return f"input_files/IC_ser5P{cfg['ser5p']}_actin{z}_RBP{cfg['rbp']}_0RNP_promoter{promoter_length}_bs{box_size}.data"
    return f"input_files/IC_ser5P{cfg['ser5p']}_RBP{cfg['rbp']}_0RNP_promoter{promoter_length}_bs{box_size}.data"
