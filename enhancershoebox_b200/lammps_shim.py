"""A `lammps`-module-shaped front end for the reference drivers' own command stream.

The reference scripts drive their engine as ``lmp = lammps(); L = IPyLammps(ptr=lmp)`` and then issue string
commands and ``gather_atoms`` (run_single_shoebox.py:323-325, 529-680, 793-899, 913-914, 1308;
VisitorGeneMaps/run_single_GENES_STAGES.py:361-363, 533-652, 738-823, 871-978).  This module offers the same two
names for exactly that command set, backed by libesb.so: ``from enhancershoebox_b200.lammps_shim import lammps,
IPyLammps`` in place of ``from lammps import lammps, IPyLammps`` runs ONE repeat chunk by chunk on the GPU with the
script's Python logic (measurements, gene state machine, ``set atom ... type ...``) left where it is.  It is the
compatibility path, not the fast one: the fused device programme (engine.ShoeboxBatch + set_schedule) steps whole
batches without the per-chunk host round trip.

Supported commands = what the two drivers issue: log, atom_style, boundary, read_data, neighbor, neigh_modify,
comm_modify, group (only `Anchors id ...` matters), angle_style cosine / angle_coeff, bond_style harmonic / bond_coeff,
pair_style soft | table lookup N, pair_coeff, variable NAME equal ramp(a,b), fix ... adapt / nve / langevin /
wall/harmonic / setforce, unfix, timestep, run N [start S stop E], set atom ID type T, and the output commands
(dump, dump_modify, undump, compute, uncompute, thermo, region), which are accepted and ignored.  Anything else raises.
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np

from . import datafile, engine as E, protocol as P


class LammpsShimError(RuntimeError):
    pass


class lammps:                                           # noqa: N801 (the reference's class name)
    def __init__(self, name="", cmdargs=None, device: int = 0, table_file: str | None = None):
        self.device, self.table_file = device, table_file
        self.data = None
        self.eng = None
        self.pair_style = None
        self.soft_stream, self.table_stream = [], []
        self.maps_dirty = True
        self.variables, self.fixes = {}, {}
        self.anchors = None
        self.seed = 1
        self.dt, self.t_target, self.t_damp = P.DT, P.LANGEVIN_T, P.LANGEVIN_DAMP
        self.wall = (P.WALL_EPS, P.WALL_CUT)
        self.skin, self.delay = P.NEIGH_SKIN, P.NEIGH_DELAY
        self.bond_coeff, self.angle_coeff = {}, {}
        self.types = None                               # host copy, changed by `set atom ID type T`
        self.types_dirty = False
        self.n_runs = 0
        self.specs = []

    # ---- the lammps module's own entry points ----------------------------------------------------
    def command(self, cmd: str):
        words = cmd.split()
        if not words:
            return
        handler = getattr(self, "_cmd_" + words[0], None)
        if handler is None:
            raise LammpsShimError(f"command not in the drivers' command set: {cmd!r}")
        handler(words[1:], cmd)

    def get_natoms(self) -> int:
        return int(self.data.n_atoms)

    def gather_atoms(self, name: str, dtype: int, count: int):
        """lmp.gather_atoms('x', 1, 3) -> ctypes doubles [3 N]; ('type', 0, 1) -> ctypes ints [N]; atom-ID order."""
        if name == "x" and dtype == 1 and count == 3:
            x = self._state()[0]
            return (C.c_double * x.size)(*x.ravel())
        if name == "v" and dtype == 1 and count == 3:
            v = self._state()[1]
            return (C.c_double * v.size)(*v.ravel())
        if name == "type" and dtype == 0 and count == 1:
            t = self.types if self.types is not None else np.asarray(self.data.types, np.int32)
            return (C.c_int * t.size)(*[int(q) for q in t])
        raise LammpsShimError(f"gather_atoms({name!r}, {dtype}, {count}) is not used by the drivers")

    def close(self):
        if self.eng is not None:
            self.eng.close()
            self.eng = None

    # ---- commands --------------------------------------------------------------------------------
    def _ignore(self, words, cmd):
        pass

    _cmd_log = _cmd_atom_style = _cmd_boundary = _cmd_comm_modify = _cmd_dump = _cmd_dump_modify = _cmd_undump = _ignore
    _cmd_compute = _cmd_uncompute = _cmd_thermo = _cmd_thermo_style = _cmd_thermo_modify = _cmd_region = _cmd_special_bonds = _ignore

    def _cmd_read_data(self, words, cmd):
        self.data = datafile.read_data(words[0])
        self.types = np.asarray(self.data.types, np.int32).copy()

    def _cmd_neighbor(self, words, cmd):
        self.skin = float(words[0])

    def _cmd_neigh_modify(self, words, cmd):
        kv = dict(zip(words[0::2], words[1::2]))
        self.delay = int(kv.get("delay", self.delay))
        if kv.get("every", "1") != "1" or kv.get("check", "yes") != "yes":
            raise LammpsShimError("only `neigh_modify every 1 delay N check yes` is implemented")

    def _cmd_group(self, words, cmd):
        if words[0] == "Anchors" and words[1] == "id":
            self.anchors = np.array([int(w) - 1 for w in words[2:]], np.int32)

    def _cmd_angle_style(self, words, cmd):
        if words[0] != "cosine":
            raise LammpsShimError("angle_style cosine only")

    def _cmd_bond_style(self, words, cmd):
        if words[0] != "harmonic":
            raise LammpsShimError("bond_style harmonic only")

    def _cmd_angle_coeff(self, words, cmd):
        self.angle_coeff[int(words[0])] = float(words[1])

    def _cmd_bond_coeff(self, words, cmd):
        self.bond_coeff[int(words[0])] = (float(words[1]), float(words[2]))

    def _cmd_pair_style(self, words, cmd):
        if words[0] == "soft":
            self.pair_style = "soft"
        elif words[0] == "table" and words[1] == "lookup":
            if int(words[2]) != P.TABLE_LENGTH:
                raise LammpsShimError(f"pair_style table lookup {P.TABLE_LENGTH} only")
            self.pair_style = "table"
            self.table_stream = []                                      # a new pair_style forgets the old coefficients
        else:
            raise LammpsShimError(f"pair_style not in the drivers' command set: {cmd!r}")
        self.maps_dirty = True

    def _cmd_pair_coeff(self, words, cmd):
        if self.pair_style == "soft":
            self.soft_stream.append((words[0], words[1], float(words[3])))     # pair_coeff I J A cutoff (A is ramped by fix adapt)
        elif self.pair_style == "table":
            self.table_stream.append((words[0], words[1], (words[3], float(words[4]))))   # I J file KEY cut
        else:
            raise LammpsShimError("pair_coeff before pair_style")
        self.maps_dirty = True

    def _cmd_variable(self, words, cmd):
        m = re.fullmatch(r"ramp\(([^,]+),([^)]+)\)", "".join(words[2:]))
        if words[1] != "equal" or not m:
            raise LammpsShimError(f"only `variable NAME equal ramp(a,b)` is implemented: {cmd!r}")
        self.variables[words[0]] = (float(m.group(1)), float(m.group(2)))

    def _cmd_fix(self, words, cmd):
        fid, style = words[0], words[2]
        if style == "adapt":                       # s1 all adapt 1 pair soft a * * v_prefactor | s2 all adapt 10 bond harmonic r0 1 v_name
            ramp = self.variables[words[-1][2:]]
            if words[4] == "pair":
                if int(words[3]) != 1 or ramp != P.SOFT_A_RAMP:
                    raise LammpsShimError(f"the engine ramps pair soft a as the drivers do (every step, ramp{P.SOFT_A_RAMP}): {cmd!r}")
                self.fixes[fid] = ("adapt_soft", 1, ramp)
            else:
                bt = int(words[7])
                if int(words[3]) != 10 or ramp != P.BOND_R0_RAMP[bt]:
                    raise LammpsShimError(f"the engine ramps bond r0 {bt} as the drivers do (every 10 steps, ramp{P.BOND_R0_RAMP[bt]}): {cmd!r}")
                self.fixes[fid] = ("adapt_bond", 10, bt, ramp)
        elif style == "nve":
            self.fixes[fid] = ("nve",)
        elif style == "langevin":
            self.t_target, self.t_damp, self.seed = float(words[3]), float(words[5]), int(words[6])
            self.fixes[fid] = ("langevin",)
        elif style == "wall/harmonic":
            self.wall = (float(words[5]), float(words[7]))          # xlo EDGE eps sigma cutoff ...
            self.fixes[fid] = ("wall",)
        elif style == "setforce":
            self.fixes[fid] = ("setforce",)
        else:
            raise LammpsShimError(f"fix style not in the drivers' command set: {cmd!r}")

    def _cmd_unfix(self, words, cmd):
        self.fixes.pop(words[0], None)

    def _cmd_timestep(self, words, cmd):
        self.dt = float(words[0])

    def _cmd_set(self, words, cmd):
        if words[0] == "atom" and words[2] == "type":
            self.types[int(words[1]) - 1] = int(words[3])
            self.types_dirty = True
        else:
            raise LammpsShimError(f"only `set atom ID type T` is implemented: {cmd!r}")

    def _cmd_run(self, words, cmd):
        n_steps = int(words[0])
        ramp = None
        if len(words) >= 5 and words[1] == "start" and words[3] == "stop":
            ramp = (int(words[2]), int(words[4]))
        self._ensure_engine()
        self._sync_types()
        pair = self._sync_pair()
        adapt = [f for f in self.fixes.values() if f[0].startswith("adapt")]
        plan = P.ChunkPlan(index=self.n_runs, pair=pair, ramp=ramp,
                           adapt_soft=any(f[0] == "adapt_soft" for f in adapt),
                           adapt_bond1=any(f[0] == "adapt_bond" and f[2] == 1 for f in adapt),
                           adapt_bond2=any(f[0] == "adapt_bond" and f[2] == 2 for f in adapt), n_steps=n_steps)
        self.eng.set_schedule(plans=[plan], observe=False)
        self.eng.run(0, 1)
        st = int(self.eng.status()[0])
        if st:
            raise LammpsShimError(f"run aborted (status {st}: " + ("lost atom / particle on or inside wall" if st & 1 else "pair distance < table inner cutoff"
                                  if st & 2 else "non-finite force") + ")")          # LAMMPS_EXCEPTIONS=ON: the driver dies, the launcher retries
        self.n_runs += 1

    # ---- engine plumbing -------------------------------------------------------------------------
    def _ensure_engine(self):
        if self.eng is not None:
            return
        if self.data is None:
            raise LammpsShimError("run before read_data")
        d = self.data
        if self.anchors is not None and not np.array_equal(np.sort(self.anchors), np.sort(d.anchors())):
            raise LammpsShimError("the Anchors group differs from the polymer ends of the data file")
        for bt, (k, r0) in self.bond_coeff.items():
            if abs(k - P.BOND_K) > 1e-12 or abs(r0 - P.BOND_R0[bt]) > 1e-12:
                raise LammpsShimError(f"bond_coeff {bt} {k} {r0}: the engine is built for the drivers' {P.BOND_K} / {P.BOND_R0[bt]}")
        for at, k in self.angle_coeff.items():
            if abs(k - P.ANGLE_K[at]) > 1e-12:
                raise LammpsShimError(f"angle_coeff {at} {k}: the engine is built for the drivers' {P.ANGLE_K[at]}")
        self.eng = E.ShoeboxBatch(d, n_replicas=1, seeds=[self.seed], device=self.device, table_file=self.table_file,
                                  neigh_skin=self.skin, neigh_delay=self.delay)
        self.specs = list(self.eng.tables.specs)

    def _state(self):
        self._ensure_engine()
        x, v, _ = self.eng.state()
        return x[0], v[0]

    def _sync_types(self):
        if self.types_dirty:
            x, v = self._state()
            self.eng.upload_state(x[None], v[None], self.types[None])
            self.types_dirty = False

    def _sync_pair(self) -> str:
        """Replay the pair_coeff stream (LAMMPS wildcards, later commands override) into the engine's pair set-up."""
        if self.pair_style == "soft":
            if self.maps_dirty:
                m = np.zeros((P.MAXT, P.MAXT))
                for (i, j), v in P._replay(self.soft_stream).items():
                    m[i, j] = m[j, i] = v
                m = np.ascontiguousarray(m)
                self.eng._check(self.eng.lib.esb_set_soft(self.eng.h, m.ctypes.data_as(C.POINTER(C.c_double))))
                self.maps_dirty = False
            return "soft"
        if self.maps_dirty:
            eff = P._replay(self.table_stream)
            new = [s for s in dict.fromkeys(eff.values()) if s not in self.specs]
            if new:
                raise LammpsShimError(f"pair tables outside the drivers' set: {new}")      # (upload them with ShoeboxBatch(condition=...) instead)
            m = np.zeros((P.MAXT, P.MAXT), np.int32)
            for (i, j), s in eff.items():
                m[i, j] = m[j, i] = self.specs.index(s)
            m = np.ascontiguousarray(m)
            self.eng._check(self.eng.lib.esb_set_table_map(self.eng.h, E.MAP_IDS["T"], m.ctypes.data_as(C.POINTER(C.c_int))))
            self.maps_dirty = False
        return "T"                                                       # the map slot this front end owns


class _System:
    def __init__(self, lmp):
        self._lmp = lmp

    @property
    def natoms(self):
        return self._lmp.get_natoms()

    def __getattr__(self, name):
        lo, hi = self._lmp.data.box_lo, self._lmp.data.box_hi
        axes = {"xlo": lo[0], "ylo": lo[1], "zlo": lo[2], "xhi": hi[0], "yhi": hi[1], "zhi": hi[2]}
        if name in axes:
            return float(axes[name])
        raise AttributeError(name)


class IPyLammps:                                        # noqa: N801
    """``L.command_name(arg, ...)`` = the command line `command_name arg ...` (PyLammps' convention)."""

    def __init__(self, ptr=None, **kw):
        self.lmp = ptr if ptr is not None else lammps(**kw)
        self.system = _System(self.lmp)

    def __getattr__(self, name):
        def issue(*args):
            self.lmp.command(" ".join([name] + [str(a) for a in args]))
        return issue
