"""ctypes front-end of the CPU oracle (oracle/esb_oracle.c) + numpy restatements of the
per-chunk measurements and of the gene state machine.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  PARITY UNPINNED for the MD part (see esb_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libesb_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (portable x86-64 code: the .so travels to the GPU box)."""
    src = os.path.join(_HERE, "esb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def build_native() -> str:
    """-march=native build for the CPU baseline (compiled on the machine that runs it); the ctypes
    handle is switched to it."""
    global _lib, _LIB_PATH
    src = os.path.join(_HERE, "esb_oracle.c")
    out = os.path.join(_HERE, "libesb_oracle_native.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src) or os.environ.get("ESB_ORACLE_REBUILD"):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", out, src, "-lm"])
    _LIB_PATH, _lib = out, None
    return out


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.esbo_create.restype = vp
        L.esbo_create.argtypes = [C.c_int, C.c_int]
        L.esbo_destroy.argtypes = [vp]
        L.esbo_set_box.argtypes = [vp, dp, dp]
        L.esbo_set_atoms.argtypes = [vp, ip, dp, dp]
        L.esbo_set_types.argtypes = [vp, ip]
        L.esbo_get_state.argtypes = [vp, dp, dp, dp, ip]
        L.esbo_set_frozen.argtypes = [vp, C.c_int, ip]
        L.esbo_set_topology.argtypes = [vp, C.c_int, ip, C.c_int, ip]
        L.esbo_get_special.argtypes = [vp, ip, ip]
        L.esbo_get_special.restype = C.c_int
        L.esbo_set_bond_coeff.argtypes = [vp, C.c_int, C.c_double, C.c_double]
        L.esbo_set_angle_coeff.argtypes = [vp, C.c_int, C.c_double]
        L.esbo_get_bond_r0.argtypes = [vp, C.c_int]
        L.esbo_get_bond_r0.restype = C.c_double
        L.esbo_get_soft_a.argtypes = [vp]
        L.esbo_get_soft_a.restype = C.c_double
        L.esbo_set_fixes.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]
        L.esbo_set_seed.argtypes = [vp, C.c_uint64, C.c_uint64]
        L.esbo_set_skin.argtypes = [vp, C.c_double]
        L.esbo_set_neigh_delay.argtypes = [vp, C.c_int]
        L.esbo_pair_soft.argtypes = [vp, C.c_double, dp]
        L.esbo_pair_table.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, dp, ip]
        L.esbo_pair_table_remap.argtypes = [vp, ip]
        L.esbo_clear_adapt.argtypes = [vp]
        L.esbo_add_adapt.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.esbo_compute_forces.argtypes = [vp, C.c_int, C.c_int, dp, dp]
        L.esbo_compute_forces.restype = C.c_int
        L.esbo_run.argtypes = [vp, C.c_int, C.c_longlong, C.c_longlong]
        L.esbo_run.restype = C.c_int
        L.esbo_get_step.argtypes = [vp]
        L.esbo_get_step.restype = C.c_longlong
        L.esbo_get_eval_counter.argtypes = [vp]
        L.esbo_get_eval_counter.restype = C.c_ulonglong
        L.esbo_get_status.argtypes = [vp]
        L.esbo_get_status.restype = C.c_int
        L.esbo_get_counters.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.esbo_reset_counters.argtypes = [vp]
        L.esbo_table_build.argtypes = [C.c_int, C.c_double, C.c_double, dp, dp, C.c_double, C.c_int, dp, dp, dp]
        L.esbo_table_build.restype = C.c_int
        L.esbo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.esbo_u32_to_unit.argtypes = [C.c_uint32]
        L.esbo_u32_to_unit.restype = C.c_double
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    o = (C.c_uint32 * 4)()
    lib().esbo_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def build_lookup_table(section, cut: float, tablength: int = 30000):
    """LOOKUP arrays of one ``pair_coeff I J file KEY cut`` -> (e[tlm1], f[tlm1], innersq, delta)."""
    tlm1 = tablength - 1
    e = np.empty(tlm1)
    f = np.empty(tlm1)
    prm = np.empty(3)
    ef = np.ascontiguousarray(section.e, dtype=np.float64)
    ff = np.ascontiguousarray(section.f, dtype=np.float64)
    rc = lib().esbo_table_build(int(section.n), float(section.rlo), float(section.rhi), _dp(ef), _dp(ff),
                                float(cut), int(tablength), _dp(e), _dp(f), _dp(prm))
    if rc != 0:
        raise ValueError("Pair table cutoff outside of table")
    return e, f, float(prm[0]), float(prm[1])


class Oracle:
    """One replica on the CPU, driven chunk by chunk with the reference's command stream."""

    def __init__(self, data, tables=None, lookup=None, seed: int = 1, skin: float = 0.3, neigh_delay: int = 10):
        """`data`: ShoeboxData.  `tables`: protocol.PairTables.  `lookup`: list of (e, f, innersq, delta)
        per table spec (see `build_lookup_table`)."""
        from enhancershoebox_b200 import protocol as P
        self.P = P
        self.L = lib()
        self.n = data.n_atoms
        self.h = self.L.esbo_create(self.n, data.n_atom_types)
        lo = np.ascontiguousarray(data.box_lo, np.float64)
        hi = np.ascontiguousarray(data.box_hi, np.float64)
        self.L.esbo_set_box(self.h, _dp(lo), _dp(hi))
        t = np.ascontiguousarray(data.types, np.int32)
        x = np.ascontiguousarray(data.x, np.float64)
        v = np.ascontiguousarray(data.v, np.float64)
        self.L.esbo_set_atoms(self.h, _ip(t), _dp(x), _dp(v))
        b = np.ascontiguousarray(data.bonds, np.int32)
        a = np.ascontiguousarray(data.angles, np.int32)
        self.L.esbo_set_topology(self.h, b.shape[0], _ip(b), a.shape[0], _ip(a))
        anchors = np.ascontiguousarray(data.anchors(), np.int32)
        self.L.esbo_set_frozen(self.h, anchors.shape[0], _ip(anchors))
        for bt in (1, 2):
            self.L.esbo_set_bond_coeff(self.h, bt, P.BOND_K, P.BOND_R0[bt])
            self.L.esbo_set_angle_coeff(self.h, bt, P.ANGLE_K[bt])
        self.L.esbo_set_fixes(self.h, P.DT, P.LANGEVIN_T, P.LANGEVIN_DAMP, 1, P.WALL_EPS, P.WALL_CUT)
        self.L.esbo_set_seed(self.h, seed, 0)
        self.L.esbo_set_skin(self.h, skin)
        self.L.esbo_set_neigh_delay(self.h, neigh_delay)          # neigh_modify every 1 delay 10 check yes (:533)
        self.tables = tables
        self.lookup = lookup
        self._pair = None
        self.soft_cut = np.ascontiguousarray(P.soft_cutoffs(), np.float64)

    def __del__(self):
        try:
            self.L.esbo_destroy(self.h)
        except Exception:
            pass

    # --- configuration ---------------------------------------------------------------------
    def set_langevin(self, mode: int):
        """1: drag + noise (reference), 2: drag only (zero-noise parity runs), 0: off."""
        P = self.P
        self.L.esbo_set_fixes(self.h, P.DT, P.LANGEVIN_T, P.LANGEVIN_DAMP, mode, P.WALL_EPS, P.WALL_CUT)

    def set_pair(self, pair: str):
        if pair == self._pair:
            return
        if pair == "soft":
            self.L.esbo_pair_soft(self.h, 0.0, _dp(self.soft_cut))
        else:
            tm = np.ascontiguousarray(self.tables.maps[pair], np.int32).copy()
            tm[tm < 0] = 0
            if self._pair in (None, "soft"):
                nt = len(self.lookup)
                tlm1 = self.lookup[0][0].shape[0]
                f = np.ascontiguousarray(np.stack([l[1] for l in self.lookup]))
                e = np.ascontiguousarray(np.stack([l[0] for l in self.lookup]))
                inner = np.array([l[2] for l in self.lookup])
                delta = np.array([l[3] for l in self.lookup])
                cut = np.array([c for (_, c) in self.tables.specs], np.float64)
                self.L.esbo_pair_table(self.h, nt, tlm1, _dp(f), _dp(e), _dp(inner), _dp(delta), _dp(cut), _ip(tm))
            else:
                self.L.esbo_pair_table_remap(self.h, _ip(tm))
        self._pair = pair

    def apply_plan(self, plan):
        """Issue the commands that precede ``run`` in chunk `plan.index`."""
        P = self.P
        self.set_pair(plan.pair)
        self.L.esbo_clear_adapt(self.h)
        if plan.adapt_soft:
            self.L.esbo_add_adapt(self.h, 0, 0, 1, *P.SOFT_A_RAMP)
        if plan.adapt_bond1:
            self.L.esbo_add_adapt(self.h, 1, 1, 10, *P.BOND_R0_RAMP[1])
        if plan.adapt_bond2:
            self.L.esbo_add_adapt(self.h, 1, 2, 10, *P.BOND_R0_RAMP[2])

    def run_chunk(self, plan) -> int:
        self.apply_plan(plan)
        if plan.ramp is not None:
            return self.L.esbo_run(self.h, plan.n_steps, plan.ramp[0], plan.ramp[1])
        return self.L.esbo_run(self.h, plan.n_steps, -1, -1)

    def run(self, nsteps: int, start: int = -1, stop: int = -1) -> int:
        return self.L.esbo_run(self.h, nsteps, start, stop)

    # --- state ------------------------------------------------------------------------------
    def state(self):
        x = np.empty((self.n, 3))
        v = np.empty((self.n, 3))
        f = np.empty((self.n, 3))
        t = np.empty(self.n, np.int32)
        self.L.esbo_get_state(self.h, _dp(x), _dp(v), _dp(f), _ip(t))
        return x, v, f, t

    def set_state(self, types, x, v=None):
        t = np.ascontiguousarray(types, np.int32)
        x = np.ascontiguousarray(x, np.float64)
        vv = None if v is None else np.ascontiguousarray(v, np.float64)
        self.L.esbo_set_atoms(self.h, _ip(t), _dp(x), None if vv is None else _dp(vv))

    def set_types(self, types):
        t = np.ascontiguousarray(types, np.int32)
        self.L.esbo_set_types(self.h, _ip(t))

    def set_seed(self, seed: int, eval_counter: int = 0):
        self.L.esbo_set_seed(self.h, seed, eval_counter)

    def forces(self, with_post: bool = False, with_noise: bool = False):
        f = np.empty((self.n, 3))
        e = np.empty(4)
        st = self.L.esbo_compute_forces(self.h, int(with_post), int(with_noise), _dp(f), _dp(e))
        return f, e, st

    def counters(self):
        out = (C.c_longlong * 5)()
        self.L.esbo_get_counters(self.h, out)
        return dict(evals=out[0], builds=out[1], pairs_in_cut=out[2], pairs_listed=out[3], dangerous=out[4])

    def reset_counters(self):
        self.L.esbo_reset_counters(self.h)

    def special(self):
        start = np.empty(self.n + 1, np.int32)
        tot = self.L.esbo_get_special(self.h, _ip(start), None)
        sp = np.empty(max(tot, 1), np.int32)
        self.L.esbo_get_special(self.h, _ip(start), _ip(sp))
        return start, sp[:tot]

    @property
    def step(self):
        return self.L.esbo_get_step(self.h)

    @property
    def status(self):
        return self.L.esbo_get_status(self.h)

    def bond_r0(self, bt):
        return self.L.esbo_get_bond_r0(self.h, bt)

    def soft_a(self):
        return self.L.esbo_get_soft_a(self.h)
