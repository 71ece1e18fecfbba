"""Restatement of the reference's actin events (TEST INFRASTRUCTURE, see esb_oracle.c).

Follows ``run_single_shoebox.py:969-1298`` statement by statement -- nucleation (:973-1038), plus/minus-end
polymerisation (:1062-1135), minus-end (:1168-1214) and plus-end (:1220-1250) depolymerisation with the reference's own list
bookkeeping (``free_actin``, ``filament_details``) -- with two substitutions, both shared with the device
(enhancershoebox_b200/csrc/esb_kernels.cu, actin_events):

* ``random.random()`` / ``np.random.uniform`` (unseeded in the reference) draw from Philox4x32-10 keyed by the
  replica seed with counter (chunk, draw index, 0, 3); one call per decision, one call (three words) per trial
  position; uniform = (word >> 8) * 2**-24;
* coordinates are the device's fp32 state: a relocated monomer's new position is rounded to fp32 when it is stored
  (``Group.L.set('atom', ...)``), and that rounded value is what later distances see.

Parity unpinned (no LAMMPS, no reference output of this path exists); what the tests pin is device == this
restatement, event by event, on the same coordinates.
"""
from __future__ import annotations

import numpy as np

from . import esb_oracle as O

SIG_ACTIN = 0.3
W_SA = 0.2
FIL_END_RADIUS = 3
MONOMER_RADIUS = 2
P_ON = 0.3
T_ON_PLUS = 1
T_ON_MINUS = 40
T_OFF_MINUS = 5
T_OFF_PLUS = 4000                       # 2 * NRuns (run_single_shoebox.py:496)
MAX_TRIES = 2000


class ActinState:
    """free_actin / filament_details of one replica (atom ids are 0-based here; the reference's are 1-based)."""

    def __init__(self, free_actin, filament_details, seed: int, p_nucleation: float):
        self.free_actin = [int(a) for a in free_actin]
        self.filament_details = [[int(a) for a in f] for f in filament_details]
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.p_nucleation = float(p_nucleation)
        self.bonds = set()          # {(a, b)} a < b, type 2
        self.angles = set()         # {(end, vertex, end)} with end < end
        for f in self.filament_details:
            for a, b in zip(f[:-1], f[1:]):
                self._bond(a, b)
            for a, b, c in zip(f[:-2], f[1:-1], f[2:]):
                self._angle(a, b, c)

    # create_bonds single/bond 2 a b ; single/angle 2 a b c
    def _bond(self, a, b):
        self.bonds.add((min(a, b), max(a, b)))

    def _angle(self, a, b, c):
        self.angles.add((min(a, c), b, max(a, c)))

    # delete_bonds GROUP atom 3 any remove special ; GROUP angle 2 any remove special  (GROUP = one atom here)
    def _delete_touching(self, atom):
        self.bonds = {b for b in self.bonds if atom not in b}
        self.angles = {g for g in self.angles if atom not in g}

    def exclusions(self, atom):
        """1-2, 1-3, 1-4 partners (LAMMPS default special_bonds lj 0 0 0, recomputed by `special yes` / `remove special`)."""
        adj = {}
        for a, b in self.bonds:
            adj.setdefault(a, set()).add(b)
            adj.setdefault(b, set()).add(a)
        seen, cur = {atom}, {atom}
        for _ in range(3):
            cur = {n for a in cur for n in adj.get(a, ()) if n not in seen}
            seen |= cur
        return seen - {atom}

    def events(self, i: int, x: np.ndarray, flags: int):
        """The actin block of Python timestep i on coordinates x [N,3] (float32 values; modified in place).
        flags as enhancershoebox_b200.protocol.actin_flags.  Returns the number of Philox calls consumed."""
        if not (flags & 1):
            return 0
        m = [0]

        def draw():
            w = O.philox((i, m[0], 0, 3), self.key)
            m[0] += 1
            return w

        def uni(w):
            return float(w >> 8) * 2.0 ** -24

        def pos(a):
            return x[a].astype(np.float64)

        def dist(p, q):
            return float(np.sqrt(np.sum((p - q) ** 2)))

        def uniform(c, h):
            w = draw()
            low, high = c - h, c + h
            return low + (high - low) * np.array([uni(w[0]), uni(w[1]), uni(w[2])])

        def place(a, q):
            x[a] = q.astype(np.float32)

        free_actin, filament_details = self.free_actin, self.filament_details
        half, lo, hi = SIG_ACTIN + W_SA, SIG_ACTIN - W_SA, SIG_ACTIN + W_SA
        # ---- nucleation (:973-1038) ----
        free_actin_positions = np.array([pos(a) for a in free_actin]).reshape(-1, 3)
        pairs_to_be_formed = []
        for j in range(len(free_actin)):
            if uni(draw()[0]) <= self.p_nucleation and i % T_ON_PLUS == 0:
                curr_atom = free_actin[j]
                other_free_actin = list.copy(free_actin)
                other_free_actin.remove(curr_atom)
                pds = [dist(pos(curr_atom), free_actin_positions[k, :]) for k in range(len(free_actin))]
                pds.remove(np.min(np.array(pds)))
                pds = np.array(pds)
                if len(pds) > 0 and np.min(pds) < MONOMER_RADIUS:
                    to_be_added = other_free_actin[int(np.argmin(pds))]
                    if any(curr_atom in sl for sl in pairs_to_be_formed) or any(to_be_added in sl for sl in pairs_to_be_formed):
                        continue
                    pairs_to_be_formed.append([curr_atom, to_be_added])
                    n_tries = 1
                    while 1:
                        n_tries = n_tries + 1
                        new_loc = uniform(pos(curr_atom), half)
                        d = dist(pos(curr_atom), new_loc)
                        if d > lo and d < hi:
                            break
                        if n_tries > MAX_TRIES:
                            break
                    if n_tries > MAX_TRIES:
                        continue
                    place(to_be_added, new_loc)
        for curr_atom, to_be_added in pairs_to_be_formed:
            self._bond(curr_atom, to_be_added)
            free_actin.remove(curr_atom)
            free_actin.remove(to_be_added)
            filament_details.append([curr_atom, to_be_added])
        # ---- polymerisation (:1062-1135) ----
        for j in range(len(filament_details)):
            if i % T_ON_PLUS == 0 and uni(draw()[0]) <= P_ON:
                self._grow(j, True, x, pos, dist, uniform, place, half, lo, hi)
            if i % (T_ON_PLUS * T_ON_MINUS) == 0 and uni(draw()[0]) <= P_ON:
                self._grow(j, False, x, pos, dist, uniform, place, half, lo, hi)
        # ---- depolymerisation at the minus end (:1168-1214) ----
        if i % T_OFF_MINUS == 0:
            assert flags & 4
            for f in filament_details:
                self._delete_touching(f[-1])
            for j in range(len(filament_details)):
                curr_atom = filament_details[j][-1]
                next_atom = filament_details[j][-2]
                free_actin.append(curr_atom)
                filament_details[j].remove(curr_atom)
                n_tries = 0
                while 1:
                    n_tries = n_tries + 1
                    new_loc = uniform(pos(curr_atom), 1.0)
                    d = dist(pos(next_atom), new_loc)
                    if d > SIG_ACTIN:
                        break
                    if n_tries > MAX_TRIES:
                        break
                if not n_tries > MAX_TRIES:
                    place(curr_atom, new_loc)
                if len(filament_details[j]) < 2:
                    free_actin.append(filament_details[j][0])
            for j in sorted([j for j in range(len(filament_details)) if len(filament_details[j]) < 2], reverse=True):
                del filament_details[j]
        # ---- depolymerisation at the plus end (:1220-1250; i % t_off_plus == 0, t_off_plus = 2 NRuns) ----
        if i % T_OFF_PLUS == 0:
            assert flags & 8
            for f in filament_details:
                self._delete_touching(f[0])
            for j in range(len(filament_details)):
                curr_atom = filament_details[j][0]
                next_atom = filament_details[j][1]
                free_actin.append(curr_atom)
                filament_details[j].remove(curr_atom)
                n_tries = 0
                while 1:
                    n_tries = n_tries + 1
                    new_loc = uniform(pos(curr_atom), 1.0)
                    d = dist(pos(next_atom), new_loc)
                    if d > SIG_ACTIN:
                        break
                    if n_tries > MAX_TRIES:
                        break
                if not n_tries > MAX_TRIES:
                    place(curr_atom, new_loc)
                if len(filament_details[j]) < 2:
                    free_actin.append(filament_details[j][-1])
            for j in sorted([j for j in range(len(filament_details)) if len(filament_details[j]) < 2], reverse=True):
                del filament_details[j]
        return m[0]

    def _grow(self, j, plus_end, x, pos, dist, uniform, place, half, lo, hi):
        free_actin, f = self.free_actin, self.filament_details[j]
        free_actin_copy = free_actin.copy()
        curr_atom, adj_atom = (f[0], f[1]) if plus_end else (f[-1], f[-2])
        pds = np.array([dist(pos(curr_atom), pos(a)) for a in free_actin_copy])
        if len(pds) > 0:
            for tba in list(np.where(pds < FIL_END_RADIUS))[0]:
                to_be_added = free_actin_copy[tba]
                n_tries = 0
                while 1:
                    n_tries = n_tries + 1
                    new_loc = uniform(pos(curr_atom), half)
                    d = dist(pos(curr_atom), new_loc)
                    da = dist(pos(adj_atom), new_loc)
                    if d > lo and d < hi and da > ((2 ** (1 / 2)) * SIG_ACTIN):
                        break
                    if n_tries > MAX_TRIES:
                        break
                if n_tries > MAX_TRIES:
                    continue
                place(to_be_added, new_loc)
                self._bond(curr_atom, to_be_added)
                self._angle(to_be_added, curr_atom, adj_atom)
                free_actin.remove(to_be_added)
                if plus_end:
                    f.insert(0, to_be_added)
                    curr_atom, adj_atom = f[0], f[1]
                else:
                    f.append(to_be_added)
                    curr_atom, adj_atom = f[-1], f[-2]

    def stats(self):
        """n_filaments, sum and sum of squares of the lengths, n_free_actin, min and max length (:1292-1297)."""
        lens = [len(f) for f in self.filament_details]
        return [len(lens), sum(lens), sum(v * v for v in lens), len(self.free_actin), min(lens) if lens else 0, max(lens) if lens else 0]
