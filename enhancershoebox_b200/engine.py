"""ctypes binding of libesb.so + the batch engine the drivers use instead of the `lammps` module.

`ShoeboxBatch` plays the role of the ``lammps()`` / ``IPyLammps`` pair in the reference drivers
(``run_single_shoebox.py:323-325``) for a whole BATCH of replicas: it is configured with the same
inputs (a ``make_input_file.py`` data file, the ``LJsoft_RSQ.table`` sections, the driver's
``pair_coeff`` stream, box/promoter/threshold/activation parameters) and advances all replicas on
one B200.  The library is required: if ``libesb.so`` is missing or no sm_100 GPU is present the
constructor raises -- there is no CPU fallback and the oracle is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import ljsoft, protocol as P
from .datafile import ShoeboxData

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESB_LIB") or os.path.join(_HERE, "libesb.so")   # ESB_LIB: perf experiments with a variant build
MAXT = 12
N_SHELL = 50
FRAME_DOUBLES = 9
MAP_IDS = {"A": 0, "B": 1, "T": 2}

ST_WALL, ST_INNER, ST_NONFINITE, ST_NEIGH = 1, 2, 4, 8


class EsbError(RuntimeError):
    pass


class ClosedForm(C.Structure):
    _fields_ = [("valid", C.c_int), ("eps_eff", C.c_double), ("s", C.c_double), ("rc", C.c_double), ("c", C.c_double)]


class ChunkDesc(C.Structure):
    _fields_ = [("chunk_index", C.c_int), ("n_steps", C.c_int), ("pair_mode", C.c_int), ("map_id", C.c_int),
                ("ramp_on", C.c_int), ("ramp_start", C.c_int64), ("ramp_stop", C.c_int64),
                ("adapt_soft", C.c_int), ("adapt_bond1", C.c_int), ("adapt_bond2", C.c_int),
                ("observe", C.c_int), ("expected_rnp", C.c_int),
                ("induce_on", C.c_int), ("activate_on", C.c_int), ("inactivate_on", C.c_int), ("microscopy", C.c_int),
                ("actin_flags", C.c_int)]


class ActinParams(C.Structure):
    _fields_ = [("first", C.c_int), ("count", C.c_int), ("p_nucleation", C.c_double), ("p_on", C.c_double),
                ("monomer_radius", C.c_double), ("end_radius", C.c_double), ("shell_lo", C.c_double), ("shell_hi", C.c_double),
                ("adj_min", C.c_double), ("free_box", C.c_double), ("free_min", C.c_double), ("max_tries", C.c_int)]


def actin_params(first: int, count: int, p_nucleation: float) -> "ActinParams":
    """The constants of run_single_shoebox.py:417-418, 428-429, 496, 1001-1004, 1074, 1186-1188, as the reference's own
    floating-point expressions."""
    return ActinParams(first, count, float(p_nucleation), P.P_ON, float(P.MONOMER_RADIUS), float(P.FIL_END_RADIUS),
                       P.SIG_ACTIN - P.W_SA, P.SIG_ACTIN + P.W_SA, (2 ** (1 / 2)) * P.SIG_ACTIN, 1.0, P.SIG_ACTIN, 2000)


_lib = None

_SIGNATURES = {
    "esb_last_error": (C.c_char_p, []),
    "esb_version": (C.c_int, []),
    "esb_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "esb_destroy": (C.c_int, [C.c_void_p]),
    "esb_set_box": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "esb_set_topology": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "esb_set_coeffs": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "esb_set_fixes": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]),
    "esb_set_neighbor": (C.c_int, [C.c_void_p, C.c_double, C.c_int]),
    "esb_set_actin": (C.c_int, [C.c_void_p, C.POINTER(ActinParams)]),
    "esb_set_microscopy": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int]),
    "esb_get_microscopy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_ushort)]),
    "esb_get_actin": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 6),
    "esb_get_actin_frames": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "esb_set_soft": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "esb_set_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ClosedForm)]),
    "esb_table_from_section": (C.c_int, [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                         C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "esb_set_table_map": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "esb_set_layout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "esb_set_replica_params": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "esb_upload_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]),
    "esb_download_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "esb_upload_state_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esb_download_state_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esb_n_pad": (C.c_int, [C.c_void_p]),
    "esb_set_schedule": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(ChunkDesc)]),
    "esb_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "esb_sync": (C.c_int, [C.c_void_p]),
    "esb_set_segments": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "esb_forces": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "esb_get_frames": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "esb_get_raw_distances": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "esb_get_frames_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "esb_get_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "esb_get_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "esb_reset_counters": (C.c_int, [C.c_void_p]),
    "esb_get_bond_r0": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "esb_get_step": (C.c_int64, [C.c_void_p]),
    "esb_device_ptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "esb_aggregate_grouped": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                        C.POINTER(C.c_double), C.c_void_p]),
    "esb_aggregate_runs": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
                                     C.POINTER(C.c_double), C.c_void_p]),
}
ABI_SYMBOLS = tuple(_SIGNATURES)


def load_library(path: str | None = None):
    """dlopen libesb.so and attach the prototypes of include/esb.h.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EsbError(f"{p} not found: build it with `python -m enhancershoebox_b200.build` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def build_lookup(section, cut: float, tablength: int = P.TABLE_LENGTH):
    """``pair_coeff I J file KEY cut`` -> (e, f, innersq, delta) via libesb's host builder."""
    lib = load_library()
    tlm1 = tablength - 1
    e, f, prm = np.empty(tlm1), np.empty(tlm1), np.empty(3)
    ef = np.ascontiguousarray(section.e, np.float64)
    ff = np.ascontiguousarray(section.f, np.float64)
    rc = lib.esb_table_from_section(int(section.n), float(section.rlo), float(section.rhi), _dp(ef), _dp(ff), float(cut),
                                    int(tablength), _dp(e), _dp(f), _dp(prm))
    if rc != 0:
        raise EsbError(lib.esb_last_error().decode())
    return e, f, float(prm[0]), float(prm[1])


_table_cache: dict = {}


def lookup_tables(tables: P.PairTables, table_file: str | None = None):
    """LOOKUP arrays + closed forms for every (KEY, cut) of `tables` (cached per process)."""
    key = (tuple(tables.specs), table_file)
    if key in _table_cache:
        return _table_cache[key]
    keys = sorted({k for (k, _) in tables.specs})
    if table_file is not None:
        sections = ljsoft.read_table_file(table_file, set(keys))
        missing = [k for k in keys if k not in sections]
        if missing:
            raise EsbError(f"{table_file}: sections not found: {missing}")
    else:
        sections = {k: ljsoft.generated_section(k) for k in keys}
    lookup = [build_lookup(sections[k], c) for (k, c) in tables.specs]
    cfs = []
    for (k, _c) in tables.specs:
        cf = ClosedForm()
        if k in ljsoft.SECTION_INDEX and not ljsoft.SECTIONS[ljsoft.SECTION_INDEX[k]][5]:
            g = ljsoft.section_geometry(k)
            cf.valid, cf.eps_eff, cf.s, cf.rc, cf.c = 1, g.eps_eff, g.s, g.rc, g.c
        cfs.append(cf)
    _table_cache[key] = (lookup, cfs)
    return lookup, cfs


class ShoeboxBatch:
    """A batch of independent shoebox replicas on one GPU."""

    def __init__(self, data, n_replicas: int | None = None, driver: str = "genes_stages", condition: str = "Control",
                 device: int = 0, table_file: str | None = None, seeds=None, thresholds=None, activations=None,
                 langevin_mode: int = 1, neigh_skin: float = P.NEIGH_SKIN, neigh_delay: int = P.NEIGH_DELAY,
                 p_nucleation: float | None = None):
        """`data`: one ShoeboxData (copied to all replicas) or a list with one per replica (same
        topology).  `thresholds` = -x values, `activations` = -a values (per mille, as on the CLI)."""
        self.lib = load_library()
        datas = [data] if isinstance(data, ShoeboxData) else list(data)
        d0 = datas[0]
        self.R = int(n_replicas if n_replicas is not None else len(datas))
        if len(datas) not in (1, self.R):
            raise ValueError("data must hold one system or one per replica")
        for d in datas[1:]:
            if d.n_atoms != d0.n_atoms or not np.array_equal(d.bonds, d0.bonds) or not np.array_equal(d.angles, d0.angles):
                raise ValueError("all replicas of a batch must share bead count and topology")
        self.data0 = d0
        self.N = d0.n_atoms
        self.driver, self.condition = driver, condition
        self.h = C.c_void_p()
        self._check(self.lib.esb_create(device, self.R, self.N, d0.n_atom_types, C.byref(self.h)))
        lo = np.ascontiguousarray(d0.box_lo, np.float64)
        hi = np.ascontiguousarray(d0.box_hi, np.float64)
        self._check(self.lib.esb_set_box(self.h, _dp(lo), _dp(hi)))
        b = np.ascontiguousarray(d0.bonds, np.int32)
        a = np.ascontiguousarray(d0.angles, np.int32)
        fr = np.ascontiguousarray(d0.anchors(), np.int32)
        self._check(self.lib.esb_set_topology(self.h, b.shape[0], _ip(b), a.shape[0], _ip(a), fr.shape[0], _ip(fr)))
        bk = np.array([0.0, P.BOND_K, P.BOND_K])
        br = np.array([0.0, P.BOND_R0[1], P.BOND_R0[2]])
        ak = np.array([0.0, P.ANGLE_K[1], P.ANGLE_K[2]])
        self._check(self.lib.esb_set_coeffs(self.h, _dp(bk), _dp(br), _dp(ak)))
        # actin polymerisation events (run_single_shoebox.py:969-1298): every condition of the actin driver but the
        # static-topology ones (:459-460)
        self.actin = None
        actin_ids = np.flatnonzero(np.asarray(d0.types) == P.ACTIN_TYPE)
        if p_nucleation is not None and len(actin_ids) and condition not in P.STATIC_ACTIN_CONDITIONS:
            first, count = int(actin_ids[0]), len(actin_ids)
            if not np.array_equal(actin_ids, np.arange(first, first + count)):
                raise ValueError("actin beads must be one contiguous block of atom ids")
            self.actin = actin_params(first, count, p_nucleation)
            self._check(self.lib.esb_set_actin(self.h, C.byref(self.actin)))
        self._check(self.lib.esb_set_fixes(self.h, P.DT, P.LANGEVIN_T, P.LANGEVIN_DAMP, langevin_mode, P.WALL_EPS, P.WALL_CUT))
        if "ESB_SKIN" not in os.environ and "ESB_DELAY" not in os.environ:          # (perf experiments override through the environment)
            self._check(self.lib.esb_set_neighbor(self.h, float(neigh_skin), int(neigh_delay)))
        sc = np.ascontiguousarray(P.soft_cutoffs(), np.float64)
        self._check(self.lib.esb_set_soft(self.h, _dp(sc)))
        # pair tables
        phases = ("A", "B", "T") if condition in ("Hexanediol", "JQ-1") else ("A", "B")
        self.tables = P.pair_tables(driver, condition, phases)
        lookup, cfs = lookup_tables(self.tables, table_file)
        nt = len(lookup)
        f = np.ascontiguousarray(np.stack([l[1] for l in lookup]))
        e = np.ascontiguousarray(np.stack([l[0] for l in lookup]))
        inner = np.array([l[2] for l in lookup])
        delta = np.array([l[3] for l in lookup])
        cut = np.array([c for (_, c) in self.tables.specs], np.float64)
        cf_arr = (ClosedForm * nt)(*cfs)
        self._check(self.lib.esb_set_tables(self.h, nt, f.shape[1], _dp(f), _dp(e), _dp(inner), _dp(delta), _dp(cut), cf_arr))
        for ph in phases:
            m = np.ascontiguousarray(self.tables.maps[ph], np.int32).copy()
            m[m < 0] = 0
            self._check(self.lib.esb_set_table_map(self.h, MAP_IDS[ph], _ip(m)))
        # layout for the fused measurements
        enh = np.ascontiguousarray(d0.enhancer_atoms(), np.int32)
        s5 = np.ascontiguousarray(d0.ser5p_atoms(), np.int32)
        cr = np.ascontiguousarray(P.CLUSTER_R, np.float64)
        self.gene_start, self.promoter_length = d0.gene_start(), d0.promoter_length()
        self.n_rbprnp = int(d0.rbp_atoms().shape[0])
        self._check(self.lib.esb_set_layout(self.h, self.gene_start, self.promoter_length, enh.shape[0], _ip(enh), s5.shape[0], _ip(s5), _dp(cr)))
        self.set_replica_params(seeds, thresholds, activations)
        # state
        if len(datas) == 1:
            self.upload_state(d0.x[None], d0.v[None], d0.types[None], replicate=True)
        else:
            self.upload_state(np.stack([d.x for d in datas]), np.stack([d.v for d in datas]), np.stack([d.types for d in datas]))
        self.n_chunks = 0
        self.plans = []

    # ------------------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise EsbError(f"libesb error {rc}: {self.lib.esb_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.esb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def set_replica_params(self, seeds=None, thresholds=None, activations=None):
        sd = None if seeds is None else np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, np.uint64), (self.R,)))
        th = None if thresholds is None else np.ascontiguousarray(np.broadcast_to(np.asarray(thresholds, np.int32), (self.R,)))
        pa = None if activations is None else np.ascontiguousarray(np.broadcast_to(np.asarray(activations, np.float64), (self.R,)) / 1000)
        self._check(self.lib.esb_set_replica_params(
            self.h,
            None if sd is None else sd.ctypes.data_as(C.POINTER(C.c_uint64)),
            None if th is None else _ip(th),
            None if pa is None else _dp(pa)))

    def upload_state(self, x, v, types, replicate: bool = False):
        x = np.ascontiguousarray(x, np.float64)
        t = np.ascontiguousarray(types, np.int32)
        vv = None if v is None else np.ascontiguousarray(v, np.float64)
        self._check(self.lib.esb_upload_state(self.h, _dp(x), None if vv is None else _dp(vv), _ip(t), int(replicate)))

    def state(self):
        """(x, v, types) of all replicas: the batch form of gather_atoms (run_single_shoebox.py:913-914)."""
        x = np.empty((self.R, self.N, 3))
        v = np.empty((self.R, self.N, 3))
        t = np.empty((self.R, self.N), np.int32)
        self._check(self.lib.esb_download_state(self.h, _dp(x), _dp(v), _ip(t)))
        return x, v, t

    def set_langevin(self, mode: int):
        self._check(self.lib.esb_set_fixes(self.h, P.DT, P.LANGEVIN_T, P.LANGEVIN_DAMP, mode, P.WALL_EPS, P.WALL_CUT))

    def set_schedule(self, n_chunks: int | None = None, observe: bool = True, plans=None, n_steps: int | None = None,
                     microscopy: bool = False):
        """Lay out the drivers' main loop (protocol.chunk_schedule) as the device programme.  microscopy: record the
        synthetic-microscopy voxel counts on every chunk with (i+1) % 10 == 0 (run_single_shoebox.py:1501)."""
        if plans is None:
            plans = P.chunk_schedule(self.driver, self.condition, n_chunks)
        self.mic_chunks = [pl.index for pl in plans if microscopy and (pl.index + 1) % P.MICROSCOPY_EVERY == 0]
        if self.mic_chunks:
            ex, ey, ez = (np.ascontiguousarray(e, np.float64) for e in P.MICROSCOPY_EDGES)
            self._check(self.lib.esb_set_microscopy(self.h, len(ex), _dp(ex), len(ey), _dp(ey), len(ez), _dp(ez), len(self.mic_chunks)))
        t_rnp_final = 2000 if self.driver == "genes_stages" else P.N_RUNS + P.TREATMENT_DURATION.get(self.condition, 0)
        slope = P.F_RNP_FINAL / (t_rnp_final - P.T_INDUCTION_ON)
        c_rnp = -slope * P.T_INDUCTION_ON
        descs = (ChunkDesc * len(plans))()
        for k, pl in enumerate(plans):
            d = descs[k]
            d.chunk_index = pl.index
            d.n_steps = pl.n_steps if n_steps is None else n_steps
            d.pair_mode = 1 if pl.pair == "soft" else 2
            d.map_id = 0 if pl.pair == "soft" else MAP_IDS[pl.pair]
            d.ramp_on = int(pl.ramp is not None)
            d.ramp_start, d.ramp_stop = pl.ramp if pl.ramp is not None else (0, 1)
            d.adapt_soft, d.adapt_bond1, d.adapt_bond2 = int(pl.adapt_soft), int(pl.adapt_bond1), int(pl.adapt_bond2)
            d.observe = int(observe)
            i = pl.index
            d.expected_rnp = int(((i + 1) * slope + c_rnp) * self.n_rbprnp) if (i + 1) >= P.T_INDUCTION_ON else -1
            t_act_off = P.N_RUNS if self.condition == "Flavopiridol" else 2 * P.N_RUNS + 10 ** 6   # :482, 885-886
            d.induce_on = int(i >= P.T_INDUCTION_ON)
            d.activate_on = int(P.T_INDUCTION_ON <= i < t_act_off)
            d.inactivate_on = int(i >= P.T_INDUCTION_ON)
            d.actin_flags = P.actin_flags(i, self.condition) if self.actin is not None else 0
            d.microscopy = 1 + self.mic_chunks.index(i) if i in self.mic_chunks else 0
        self._check(self.lib.esb_set_schedule(self.h, len(plans), descs))
        self.n_chunks = len(plans)
        self.plans = list(plans)

    def run(self, chunk_begin: int = 0, n_chunks: int | None = None, stream=None, sync: bool = True):
        """``L.run(tRun)`` for n_chunks Python timesteps, measurements and state machine fused."""
        n = self.n_chunks - chunk_begin if n_chunks is None else n_chunks
        self._check(self.lib.esb_run(self.h, chunk_begin, n, stream))
        if sync:
            self._check(self.lib.esb_sync(self.h))

    def sync(self):
        self._check(self.lib.esb_sync(self.h))

    def set_segments(self, n_seg: int = 0, co_scheduled: bool = False):
        """Scheduling hint for a batch whose launches share the GPU with another batch's (esb.h: esb_set_segments)."""
        self._check(self.lib.esb_set_segments(self.h, int(n_seg), int(co_scheduled)))

    def forces(self, pair: str = "B", soft_a: float = 0.0, with_post: bool = False, eval_index: int = 0, use_arrays: bool = False):
        """Parity hook: forces [R,N,3] and energies [R,4] (pair, bond, angle, wall) at the current state."""
        f = np.empty((self.R, self.N, 3))
        e = np.empty((self.R, 4))
        mode = 0 if pair is None else (1 if pair == "soft" else 2)
        mid = 0 if mode != 2 else MAP_IDS[pair]
        self._check(self.lib.esb_forces(self.h, mode, mid, float(soft_a), int(with_post), int(eval_index), int(use_arrays), _dp(f), _dp(e)))
        return f, e

    def frames(self, chunk_begin: int = 0, n_chunks: int | None = None, with_hist: bool = False):
        n = self.n_chunks - chunk_begin if n_chunks is None else n_chunks
        out = np.empty((self.R, n, FRAME_DOUBLES))
        hist = np.empty((self.R, n, N_SHELL - 1), np.int32) if with_hist else None
        self._check(self.lib.esb_get_frames(self.h, chunk_begin, n, _dp(out), None if hist is None else _ip(hist)))
        return (out, hist) if with_hist else out

    def raw_distances(self, chunk_begin: int = 0, n_chunks: int | None = None):
        """(d_rp, d_rg) in bead units, [R, n, 2]."""
        n = self.n_chunks - chunk_begin if n_chunks is None else n_chunks
        out = np.empty((self.R, n, 2))
        self._check(self.lib.esb_get_raw_distances(self.h, chunk_begin, n, _dp(out)))
        return out

    def actin_state(self):
        """Filament bookkeeping per replica: dict(plus, minus [R, count] neighbour atom ids (-1 none), free = list of
        free_actin lists, filaments = list of filament_details lists (plus end first))."""
        n = self.actin.count
        plus, minus = np.empty((self.R, n), np.int32), np.empty((self.R, n), np.int32)
        fl, fp = np.empty((self.R, n), np.int32), np.empty((self.R, n), np.int32)
        nfree, nfil = np.empty(self.R, np.int32), np.empty(self.R, np.int32)
        self._check(self.lib.esb_get_actin(self.h, _ip(plus), _ip(minus), _ip(nfree), _ip(fl), _ip(nfil), _ip(fp)))
        first = self.actin.first
        free = [list(fl[r, :nfree[r]]) for r in range(self.R)]
        fils = []
        for r in range(self.R):
            out = []
            for b in fp[r, :nfil[r]]:
                chain = [int(b)]
                while minus[r, chain[-1] - first] >= 0:
                    chain.append(int(minus[r, chain[-1] - first]))
                out.append(chain)
            fils.append(out)
        return dict(plus=plus, minus=minus, free=free, filaments=fils)

    def actin_frames(self, chunk_begin: int = 0, n_chunks: int | None = None):
        """(stats [R, n, 6] = n_filaments, sum len, sum len^2, n_free, min len, max len; hist [R, n, 49])."""
        n = self.n_chunks - chunk_begin if n_chunks is None else n_chunks
        stats = np.empty((self.R, n, 6), np.int32)
        hist = np.empty((self.R, n, 49), np.int32)
        self._check(self.lib.esb_get_actin_frames(self.h, chunk_begin, n, _ip(stats), _ip(hist)))
        return stats, hist

    def microscopy(self):
        """Voxel counts [R, frames, 5, nx, ny, nz] of the chunks in self.mic_chunks; channels Ser5P, Ser2P, chromatin,
        gene, actin (np.histogramdd of run_single_shoebox.py:157-205)."""
        shape = tuple(len(e) - 1 for e in P.MICROSCOPY_EDGES)
        out = np.empty((self.R, len(self.mic_chunks), 5) + shape, np.uint16)
        self._check(self.lib.esb_get_microscopy(self.h, 0, len(self.mic_chunks), out.ctypes.data_as(C.POINTER(C.c_ushort))))
        return out

    def status(self):
        st = np.empty(self.R, np.int32)
        self._check(self.lib.esb_get_status(self.h, _ip(st)))
        return st

    def counters(self):
        out = np.empty((self.R, 12), np.int64)
        self._check(self.lib.esb_get_counters(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(evals=out[:, 0], builds=out[:, 1], pairs=out[:, 2], listed=out[:, 3], steps=out[:, 4], total_active=out[:, 5],
                    cyc_atom=out[:, 6], cyc_build=out[:, 7], cyc_force=out[:, 8], cyc_other=out[:, 9], dangerous=out[:, 10],
                    max_pair_force=out[:, 11] / 256.0)

    def reset_counters(self):
        self._check(self.lib.esb_reset_counters(self.h))

    def bond_r0(self):
        r0 = np.empty(3)
        self._check(self.lib.esb_get_bond_r0(self.h, _dp(r0)))
        return r0

    @property
    def step(self) -> int:
        return int(self.lib.esb_get_step(self.h))

    @property
    def n_pad(self) -> int:
        return int(self.lib.esb_n_pad(self.h))

    def device_ptr(self, which: int) -> int:
        return int(self.lib.esb_device_ptr(self.h, which) or 0)


def aggregate_all(frames: np.ndarray, contact_dist: float = 250.0, t_skip: int = 150, device: int = 0):
    """dist_genestages_all.py: one row per run (VGM/dist_genestages_all.py:66-85) = groups of one."""
    return aggregate_grouped(frames, 1, contact_dist, t_skip, device)


def aggregate_grouped(frames: np.ndarray, group: int = 5, contact_dist: float = 250.0, t_skip: int = 150, device: int = 0):
    """dist_genestages_grouped.py rows (S5PInt, S2PInt, Contact, DistActivation) for frames
    [n_runs, n_frames, 9]; groups of `group` consecutive runs (VGM/dist_genestages_grouped.py:70-106)."""
    lib = load_library()
    fr = np.ascontiguousarray(frames, np.float64)
    n_runs, n_frames = fr.shape[0], fr.shape[1]
    out = np.empty((n_runs // group, 4))
    rc = lib.esb_aggregate_grouped(device, fr.ctypes.data_as(C.c_void_p), 0, n_runs, n_frames, group, contact_dist, t_skip, _dp(out), None)
    if rc != 0:
        raise EsbError(f"esb_aggregate_grouped failed with {rc}")
    return out


def _linear_percentile_index(n: int, q: float):
    """Lower order statistic and weight of numpy's default ('linear') percentile for n samples: virtual index
    (n - 1) * (q / 100), the expression of numpy.lib._function_base_impl._QuantileMethods['linear']."""
    quant = np.true_divide(q, 100)
    vi = (n - 1) * quant
    prev = int(np.floor(vi))
    return prev, float(vi - prev)


def aggregate_runs(frames: np.ndarray, q: float = 5.0, contact_dist: float = 250.0, t_skip: int = 150, device: int = 0):
    """Per-run rows [n_runs, 6]: S5PInt, S2PInt, Contact, DistActivation, q-percentile of d_rg[t_skip:] and of
    d_rp[t_skip:] (VGM/dist_genestages_grouped_5percentile_10xaveraging.py:78-89)."""
    lib = load_library()
    fr = np.ascontiguousarray(frames, np.float64)
    n_runs, n_frames = fr.shape[0], fr.shape[1]
    prev, gamma = _linear_percentile_index(n_frames - t_skip, q)
    out = np.empty((n_runs, 6))
    rc = lib.esb_aggregate_runs(device, fr.ctypes.data_as(C.c_void_p), 0, n_runs, n_frames, contact_dist, t_skip, prev, gamma, _dp(out), None)
    if rc != 0:
        raise EsbError(f"esb_aggregate_runs failed with {rc}")
    return out


def aggregate_percentile_10x(frames: np.ndarray, group: int = 10, **kw):
    """Rows of dist_genestages_grouped_5percentile_10xaveraging.py: np.nanmean of the per-run statistics over
    consecutive groups of 10 runs, plus one row for the remainder (:94-141).  Returns (labels, rows[., 6])."""
    import warnings
    per_run = aggregate_runs(frames, **kw)
    labels, rows = [], []
    n = per_run.shape[0]
    for lo in range(0, n, group):
        hi = min(lo + group, n)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rows.append([np.nanmean(list(per_run[lo:hi, c])) for c in range(6)])
        labels.append(f"{lo + 1}-{hi}")
    return labels, np.array(rows).reshape(-1, 6)
