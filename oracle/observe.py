"""numpy/scipy restatement of what the reference drivers do between two ``run`` commands:
per-chunk measurements, the gene induction/activation/inactivation state machine, the RBP->RNP
ramp, the geneTrack record and the Ser5P shell histogram.

TEST INFRASTRUCTURE (see esb_oracle.c).  Follows, expression by expression,
``run_single_shoebox.py:913-954, 956-961, 1304-1400, 1457-1519`` ==
``VisitorGeneMaps/run_single_GENES_STAGES.py:822-865, 871-978, 1033-1089``.

The reference draws from Python's unseeded ``random``; here the two random decisions use the
same counter-based Philox draws as the device (key = replica seed):
  * activation:  ctr = (chunk, 0, 0, 1) -> u in [0,1) with 24 bits; skipped when u > p
  * RBP->RNP:    ctr = (chunk, m, 0, 2) for the m-th conversion of the chunk -> index
                 ((bits >> 8) * n_rbp) >> 24 into the current RBP beads in ascending atom order
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.spatial.distance import cdist

from . import esb_oracle as O

T_INDUCED, T_ACTIVE, T_INACTIVE = 6, 7, 2
T_PROMOTER, T_ACTIVE_PROMOTER = 10, 11
T_RBP, T_RNP = 4, 5
LENGTH_GENE = 5


@dataclass
class Layout:
    """Index sets fixed at start-up (run_single_shoebox.py:683-704); all 0-based."""

    gene_start: int
    promoter_length: int
    enhancer: np.ndarray
    ser5p: np.ndarray
    n_rbprnp: int

    @classmethod
    def from_data(cls, d):
        return cls(d.gene_start(), d.promoter_length(), d.enhancer_atoms(), d.ser5p_atoms(), int(d.rbp_atoms().shape[0]))

    @property
    def probe(self) -> int:
        # atomPositions[[x-promoter_length+1 ...]] with x the 1-based gene-start ID (:941)
        return self.gene_start + 1 - self.promoter_length + 1

    @property
    def promoter_rows(self):
        # atomPositions[[y-1 for y in [gene_start_atoms[0]-1-x for x in range(promoter_length)]]] (:946)
        g1 = self.gene_start + 1
        return [g1 - 1 - k - 1 for k in range(self.promoter_length)]


def measure(x: np.ndarray, lay: Layout, radius: float = 3.0):
    """d_rg, d_rp, probe position, Ser5P counts (:932-954)."""
    RR = np.mean(x[lay.enhancer], axis=0)
    gene = x[[lay.gene_start]]
    d_rg = float(np.sqrt(np.sum((RR - gene[0, :]) ** 2)))
    probe = x[[lay.probe], :]
    d_rp = float(np.sqrt(np.sum((RR - np.mean(x[lay.promoter_rows, :], axis=0)) ** 2)))
    s5 = x[lay.ser5p, :]
    n_prom = int(np.sum(cdist(probe, s5, "euclidean")[0] < radius))
    n_gene = int(np.sum(cdist(gene, s5, "euclidean")[0] < radius))
    return dict(probe=probe[0].copy(), d_rp=d_rp, d_rg=d_rg, n_prom=n_prom, n_gene=n_gene)


def shell_histogram(x: np.ndarray, lay: Layout, cluster_r: np.ndarray):
    """49-bin Ser5P-around-enhancer histogram (:1504-1519)."""
    ds = cdist(x[lay.ser5p, :], x[lay.enhancer, :], "euclidean")
    nums = [int(np.sum((ds < cr).any(axis=1))) for cr in cluster_r]
    return [b - a for a, b in zip(nums[0:], nums[1:])]


def expected_rnp_table(n_chunks: int, n_rbprnp: int, t_rnp_final: int, t_induction_on: int = 150, f_final: float = 0.9):
    """expected_rnp per chunk (:523-526, :1457-1460); -1 where the ramp is not evaluated."""
    slope = f_final / (t_rnp_final - t_induction_on)
    c = -slope * t_induction_on
    out = np.full(n_chunks, -1, np.int32)
    for i in range(n_chunks):
        if (i + 1) >= t_induction_on:
            out[i] = int(((i + 1) * slope + c) * n_rbprnp)
    return out


@dataclass
class GeneMachine:
    """One gene's bookkeeping across chunks; mutates `types` in place like ``set atom``."""

    lay: Layout
    threshold: int            # ser5p_to_activate (-x)
    p_activation: float       # -a / 1000
    seed: int
    t_induction_on: int = 150
    activation_delay: int = 2
    t_active_duration: int = 50
    t_activation_off: int = 10 ** 9   # Flavopiridol sets it to NRuns (run_single_shoebox.py:885-886)
    state: int = 0            # 0 inactive, 1 induced, 2 active(-listed)
    marked: bool = False
    delay: int = 0
    duration: int = 0
    total_active: int = 0
    key: tuple = field(init=False)

    def __post_init__(self):
        self.key = (self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF)

    def step(self, i: int, types: np.ndarray, n_prom: int, expected_rnp: int) -> int:
        """Everything between the measurements and the geneTrack write for chunk i.
        Returns the state column of the record."""
        g, p = self.lay.gene_start, self.lay.promoter_length
        if self.state == 2:                                   # :956-961
            self.duration += 1
            self.total_active += 1
        if i >= self.t_induction_on and self.state == 0:      # :1304-1312
            types[g:g + LENGTH_GENE] = T_INDUCED
            self.state = 1
        if self.t_induction_on <= i < self.t_activation_off and self.state == 1:      # :1322-1345
            if n_prom > self.threshold:
                u = O.lib().esbo_u32_to_unit(O.philox((i, 0, 0, 1), self.key)[0])
                if not (u > self.p_activation):
                    self.marked, self.delay, self.state = True, 0, 2
        if self.marked:                                       # :1356-1368
            if self.delay < self.activation_delay:
                self.delay += 1
            else:
                self.marked = False
                types[g:g + LENGTH_GENE] = T_ACTIVE
                types[g - p:g] = T_ACTIVE_PROMOTER
        if i >= self.t_induction_on and self.state == 2 and self.duration >= self.t_active_duration:  # :1380-1392
            self.duration = 0
            types[g:g + LENGTH_GENE] = T_INACTIVE
            types[g - p:g] = T_PROMOTER
            self.state = 0
        if expected_rnp >= 0:                                 # :1457-1470
            actual = int(np.sum(types == T_RNP))
            m = 0
            while expected_rnp > actual:
                rbp = np.nonzero(types == T_RBP)[0]
                bits = O.philox((i, m, 0, 2), self.key)[0]
                k = ((bits >> 8) * int(rbp.shape[0])) >> 24
                types[rbp[k]] = T_RNP
                actual += 1
                m += 1
        return self.state


def frame_record(i: int, meas: dict, state: int, delt: float = 6.0, sig_nm: float = 60):
    """The 9 geneTrack columns as floats/ints (:1496)."""
    return [(i + 1) * delt, meas["probe"][0] * sig_nm, meas["probe"][1] * sig_nm, meas["probe"][2] * sig_nm,
            meas["d_rp"] * sig_nm, meas["d_rg"] * sig_nm, meas["n_prom"], meas["n_gene"], state]


def frame_line(rec) -> str:
    """geneTrack.txt line: ``str()`` of each column joined by commas (:1496)."""
    return ",".join(str(v) for v in rec) + "\n"


def microscopy_histograms(x: np.ndarray, types: np.ndarray) -> np.ndarray:
    """The five np.histogramdd voxel counts of makeFakeMicroscopyImages (run_single_shoebox.py:152-205), [5, 21, 17, 13]."""
    from enhancershoebox_b200 import protocol as P
    out = []
    for _name, tset in P.MICROSCOPY_CHANNELS:
        atoms = np.array([j for j in range(len(types)) if types[j] in tset], dtype=int)
        positions = np.asarray(x, np.float64)[atoms, :].reshape(-1, 3)
        H, _ = np.histogramdd(positions, bins=P.MICROSCOPY_EDGES)
        out.append(H)
    return np.array(out)
