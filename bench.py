#!/usr/bin/env python
"""Benchmark of the EnhancerShoeBox hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]

A "step" is one Python timestep of the reference drivers ("chunk"): `run 200` + the per-chunk
measurements / gene state machine / RNP ramp, over a whole batch of replicas.  The metric is
BASELINE.json's: aggregate replica-timesteps/s (bead-steps/s is reported next to it).

Workloads (config.workload):
  box12_p3_x512   box 12, promoter 3, 512 replicas per GPU  -- the per-GPU share of BASELINE.json
                  configs[3] (4096-replica box-12 ensemble on 8 GPUs); weak scaling.  DEFAULT.
  box11_p3_x444   box 11, promoter 3, 444 replicas per GPU
All replicas start from an equilibrated mid-run frame (chunk 1000 of 2000, 41 % RNP; produced by the
oracle, tests/golden/snap_box12_chunk1000.npz) with distinct Langevin seeds, so the pair counts are
those of a production run, not of the dilute initial condition.

--impl reference times the CPU restatement of the reference's LAMMPS path (oracle/, "port": LAMMPS
itself is not installable here) on all host cores, on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
T_RUN = 200
START_CHUNK = 1000

WORKLOADS = {
    "box12_p3_x512": dict(box=12, promoter=3, replicas=512, snapshot="snap_box12_chunk1000.npz", tag="1000"),
    "box11_p3_x444": dict(box=11, promoter=3, replicas=444, snapshot="snap_box11.npz", tag="500"),
}


def load_snapshot(w):
    from enhancershoebox_b200 import datafile
    z = np.load(os.path.join(ROOT, "tests", "golden", w["snapshot"]))
    d = datafile.make_shoebox(w["box"], w["promoter"], False, seed=7)
    tag = w["tag"]
    return d, z["x" + tag].astype(np.float64), z["v" + tag].astype(np.float64), z["t" + tag].astype(np.int32)


def config_of(args, w, d, world, n_pad):
    """The workload description both arms print (same keys, same values: the driver compares them)."""
    from enhancershoebox_b200 import protocol as P
    R = args.replicas or w["replicas"]
    return {"workload": args.workload, "replicas_per_gpu": R, "n_atoms": d.n_atoms, "md_steps_per_step": T_RUN,
            "start_chunk": START_CHUNK,
            "l2": f"not flushed: the K timed steps of the native arm are one persistent launch and nothing is re-read across steps "
                  f"(a step reads the {2 * R * n_pad * 16 / 1e6:.0f} MB state once, keeps it in shared memory for 200 MD steps and streams its "
                  f"own half neighbour lists, ~45 KB per replica and build, from L2)",
            "neighbor": f"skin {P.NEIGH_SKIN} delay {P.NEIGH_DELAY} check yes (reference: run_single_shoebox.py:532-533)",
            "parallelism": f"replicas sharded over {world} GPU(s), no collective on the step path"}


def smem_bytes_per_replica_step(pairs_in_cut: float, pairs_rejected: float) -> float:
    """Algorithmic shared-memory traffic: 28 B per in-cutoff pair (16 B partner read + 12 B partner force
    update), 16 B per listed pair rejected by the cutoff; every unordered pair once."""
    return 28.0 * pairs_in_cut + 16.0 * pairs_rejected


def flops_per_replica_step(pairs_in_cut: float, n_atoms: int, n_bonds: int, n_angles: int) -> float:
    """Algorithmic work: 34 flop per in-cutoff pair (each unordered pair once),
    40 per bead-step, 25 per bond, 55 per angle.  Rejected candidates are not algorithmic work."""
    return 34.0 * pairs_in_cut + 40.0 * n_atoms + 25.0 * n_bonds + 55.0 * n_angles


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons}


def bind_to_gpu_cpus(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU (same NUMA node as the PCIe root): the pinned-buffer
    copies of the end-to-end leg then stay on that node.  Returns what was done, for the record."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} CPUs local to GPU {index} ({allowed[0]}-{allowed[-1]})"
    except Exception as e:                        # no NVML / not permitted: leave the scheduler alone
        return f"not bound ({type(e).__name__})"
    return "not bound"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def latest_traffic():
    """dram bytes per launch of the step kernel from the committed ncu summary (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "latest.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------------
def cpu_reference(w, seconds_budget: float, threads: int | None = None):
    """The oracle (CPU restatement of the LAMMPS path) on `threads` host cores, one replica per
    thread, on a bounded sample of the workload: every thread advances its replica chunk by chunk
    from the same mid-run frame until the budget is spent."""
    from concurrent.futures import ThreadPoolExecutor
    from enhancershoebox_b200 import ljsoft, protocol as P
    from oracle import esb_oracle as O
    O.build_native()
    d, x, v, t = load_snapshot(w)
    pt = P.pair_tables("genes_stages")
    secs = {k: ljsoft.generated_section(k) for k in sorted({s[0] for s in pt.specs})}
    lookup = [O.build_lookup_table(secs[k], c) for (k, c) in pt.specs]
    threads = threads or (os.cpu_count() or 1)
    plan = P.chunk_schedule("genes_stages", n_chunks=START_CHUNK + 1)[START_CHUNK]

    def make(k):
        o = O.Oracle(d, pt, lookup, seed=1 + k, shared_tables=True)
        o.set_state(t, x, v)
        return o

    oracles = [make(k) for k in range(threads)]
    deadline = [0.0]

    def work(o):
        n = 0
        while True:
            if o.run_chunk(plan) != 0:
                raise RuntimeError("oracle replica failed")
            n += 1
            if time.perf_counter() >= deadline[0]:
                return n

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda o: (o.apply_plan(plan), o.run(20)), oracles))          # warm-up (neighbour lists, caches)
        t0 = time.perf_counter()
        deadline[0] = t0 + seconds_budget
        chunks = list(ex.map(work, oracles))
        dt = time.perf_counter() - t0
    steps = sum(chunks) * T_RUN
    pairs = np.mean([o.counters()["pairs_in_cut"] / max(o.counters()["evals"], 1) for o in oracles])
    return dict(value=steps / dt, unit="replica-steps/s", cores=threads, kind="port",
                sample=f"{sum(chunks)} chunks x {T_RUN} steps on {threads} threads (one replica per thread, shared read-only tables) "
                       f"from the chunk-{START_CHUNK} frame in {dt:.1f} s",
                per_thread=steps / dt / threads,
                pairs_in_cut=float(pairs), seconds=dt, bead_steps_per_s=steps / dt * d.n_atoms)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    secs = min(60.0, max(5.0, 4.0 * args.steps))
    cb = cpu_reference(w, secs)
    d, *_ = load_snapshot(w)
    n_pad = (d.n_atoms + 31) // 32 * 32
    line = {
        "impl": "reference", "metric": "replica_timesteps_per_s", "value": cb["value"], "unit": "replica-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # one step of this workload = T_RUN MD steps of every replica of the batch: what that would take at the measured rate
        "ms_per_step": 1e3 * (args.replicas or w["replicas"]) * T_RUN / cb["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, w, d, max(args.gpus, 1), n_pad),
        "note": "CPU restatement of the reference's LAMMPS command stream (kind: port -- LAMMPS itself is not installable here nor on "
                "the GPU box: profiles/r2_lammps_probe.log); ms_per_step = time for one 200-step chunk of the whole batch at the sampled rate",
        "bead_steps_per_s": cb["bead_steps_per_s"],
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "per_thread")},
        "e2e": {"value": cb["value"], "unit": "replica-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def run_native(args, w):
    import torch
    import torch.distributed as dist
    from enhancershoebox_b200 import build as B, engine as E, protocol as P
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_cpus(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if local == 0:
        B.build()                                  # no-op when libesb.so is newer than its sources; never by two ranks at once
    if world > 1:
        dist.barrier()
    d, x, v, t = load_snapshot(w)
    R = args.replicas or w["replicas"]
    K, W = args.steps, args.warmup
    eng = E.ShoeboxBatch(d, n_replicas=R, seeds=np.arange(R, dtype=np.uint64) + 1 + 100003 * rank,
                         thresholds=80, activations=30, device=local)
    eng.upload_state(x[None], v[None], t[None], replicate=True)
    plans = P.chunk_schedule("genes_stages", n_chunks=START_CHUNK + W + 2 * K + 2)[START_CHUNK:]
    eng.set_schedule(plans=plans, observe=True)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    eng.run(0, W, stream=sp, sync=False)
    barrier()
    eng.reset_counters()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.run(W, K, stream=sp, sync=False)          # ONE launch for the K timed steps: the persistent kernel hands out
    e1.record(stream)                             # (replica, chunk) items, so replicas advance in a ring over the SMs
    barrier()
    ms = e0.elapsed_time(e1)
    cnt = eng.counters()
    status = eng.status()

    # ---- end-to-end through the host-facing API: host state in, chunk, host state + records out, EVERY step ----
    # The batch runs as two half batches (two handles, two streams): while one half steps, the other half's state
    # travels over PCIe.  Every step uploads each half's state from pinned host memory, runs one chunk, downloads the
    # state (gather_atoms for the whole batch) and the chunk's geneTrack records; a half's step c + 1 is enqueued as soon
    # as the host has seen its step c - 1 (the host consumes results one step behind the device, it never skips one).
    eng.close()
    halves = []
    for hh in range(2):
        Rh = R // 2 + (R % 2 if hh == 0 else 0)
        eh = E.ShoeboxBatch(d, n_replicas=Rh, seeds=np.arange(Rh, dtype=np.uint64) + 1 + 100003 * rank + 7919 * hh,
                            thresholds=80, activations=30, device=local)
        eh.upload_state(x[None], v[None], t[None], replicate=True)
        eh.set_schedule(plans=plans, observe=True)
        eh.set_segments(8, co_scheduled=True)
        st_h = torch.cuda.Stream()
        hp = torch.empty((Rh, eh.n_pad, 4), dtype=torch.float32).pin_memory()
        hv = torch.empty((Rh, eh.n_pad, 4), dtype=torch.float32).pin_memory()
        hf = torch.empty((2, Rh, 9), dtype=torch.float64).pin_memory()
        eh.lib.esb_download_state_f32(eh.h, hp.data_ptr(), hv.data_ptr(), st_h.cuda_stream)
        halves.append(dict(eng=eh, stream=st_h, pos=hp, vel=hv, frames=hf, events={}, R=Rh))
    torch.cuda.synchronize()
    n_pad = halves[0]["eng"].n_pad
    records_seen = [0]

    def e2e_enqueue(c):
        for hb in halves:
            eh, sp_h = hb["eng"], hb["stream"].cuda_stream
            eh.lib.esb_upload_state_f32(eh.h, hb["pos"].data_ptr(), hb["vel"].data_ptr(), sp_h)        # H2D: this step's input
            eh.run(c, 1, stream=sp_h, sync=False)
            eh.lib.esb_download_state_f32(eh.h, hb["pos"].data_ptr(), hb["vel"].data_ptr(), sp_h)      # D2H: gather_atoms
            eh.lib.esb_get_frames_async(eh.h, c, 1, hb["frames"][c & 1].data_ptr(), sp_h)                # D2H: this chunk's geneTrack records
            ev = torch.cuda.Event()
            ev.record(hb["stream"])
            hb["events"][c] = ev

    def e2e_consume(c):
        for hb in halves:
            hb["events"].pop(c).synchronize()                                                           # the host reads step c's result
            records_seen[0] += int(hb["frames"][c & 1][:, 0].numel())

    first = W + K
    e2e_enqueue(first)
    e2e_consume(first)
    barrier()
    t0 = time.perf_counter()                      # end to end is what the host sees: wall clock from the first enqueue to the last result
    for c in range(first + 1, first + 1 + K):
        e2e_enqueue(c)
        if c > first + 1:
            e2e_consume(c - 1)
    e2e_consume(first + K)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)
    status_e2e = np.concatenate([hb["eng"].status() for hb in halves])
    for hb in halves:
        hb["eng"].close()

    t_max = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    fails = torch.tensor([int((status != 0).sum()) + int((status_e2e != 0).sum())], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(fails, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t_max[0]), float(t_max[1])

    if rank == 0:
        total_R = R * world
        value = total_R * T_RUN * K / (ms * 1e-3)
        e2e_value = total_R * T_RUN * K / (ms_e2e * 1e-3)
        evals = float(cnt["evals"].sum())
        pairs_in_cut = float(cnt["pairs"].sum()) / evals / 2.0
        listed = float(cnt["listed"].sum()) / evals / 2.0
        fl = flops_per_replica_step(pairs_in_cut, d.n_atoms, d.bonds.shape[0], d.angles.shape[0])
        sm_bytes = smem_bytes_per_replica_step(pairs_in_cut, listed - pairs_in_cut)
        peaks, which = measured_peaks()
        f_clk = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
        fp32_peak = 148 * 128 * 2 * f_clk / 1e12                                                # TFLOP/s
        smem_peak = 148 * 128 * f_clk / 1e9                                                     # GB/s: 128 B per clock and SM
        rate = R * T_RUN * K / (ms * 1e-3)                                                      # replica-steps/s per GPU (one launch = K steps)
        prof = latest_traffic()
        cyc = {k: float(cnt[k].sum()) for k in ("cyc_atom", "cyc_build", "cyc_force", "cyc_other")}
        cyc_tot = sum(cyc.values()) or 1.0
        force_share = cyc["cyc_force"] / cyc_tot
        frac_smem, frac_fp32 = rate * sm_bytes / 1e9 / smem_peak, rate * fl / 1e12 / fp32_peak
        line = {
            "metric": "replica_timesteps_per_s", "value": value, "unit": "replica-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, w, d, world, n_pad),
            "bead_steps_per_s": value * d.n_atoms,
            "e2e": {"value": e2e_value, "unit": "replica-steps/s", "h2d_bytes_per_step": int(2 * R * n_pad * 16),
                    "d2h_bytes_per_step": int(2 * R * n_pad * 16 + R * 72), "ms_per_step": ms_e2e / K,
                    "how": "two half batches on two streams (copies of one half overlap the steps of the other); every step: pinned host "
                           "state -> device, one chunk, state + geneTrack records -> pinned host; the host reads every step's records, one step "
                           "behind the device", "records_read": records_seen[0], "cpu_affinity": affinity},
            "gpu_launches": 1,
            # The step path is neither HBM- nor tensor-bound (state in shared memory, no dense contraction): SURVEY 8d's two
            # ceilings are the FP32 issue rate and the shared-memory bandwidth of the pair loop; the binding one is the
            # one whose ceiling costs more time per replica-step -- shared memory (531 KB vs 568 kflop per replica-step).
            "roofline": {"bound": "smem", "achieved": rate * sm_bytes / 1e9, "peak": smem_peak, "unit": "GB/s", "frac": frac_smem,
                         "frac_pair_loop": frac_smem / max(force_share, 1e-9),
                         "fp32": {"achieved": rate * fl / 1e12, "peak": fp32_peak, "unit": "TFLOP/s", "frac": frac_fp32,
                                  "frac_pair_loop": rate * 34.0 * pairs_in_cut / 1e12 / fp32_peak / max(force_share, 1e-9)},
                         "traffic": None if not prof else prof.get("dram_bytes_per_step", 0) * K * (R / max(prof.get("replicas", R), 1)),
                         "traffic_source": None if not prof else prof.get("source"),
                         # from the committed ncu capture of this kernel (profiles/latest.json), not measured live
                         "ncu_committed_capture": None if not prof else {k: prof.get(k) for k in ("ipc", "issue_slot_utilization", "warp_instructions_per_replica_step",
                                                                                "achieved_occupancy_warps_of_64", "smem_wavefronts_pct_of_peak", "l2_hit_rate_pct")},
                         "peak_source": f"148 SM x 128 B/clk (x 128 lanes x 2 flop for FP32) x sm_max_mhz ({which} MEASURED_PEAKS.json; it has no SMEM/FP32 entry)",
                         "kernel": "esb_run_kernel", "launch_ms": ms, "steps_per_launch": K, "flop_per_replica_step": fl,
                         "smem_bytes_per_replica_step": sm_bytes,
                         "pairs_in_cut_per_step": pairs_in_cut, "pairs_listed_per_step": listed,
                         "list_builds_per_step": float(cnt["builds"].sum()) / max(float(cnt["steps"].sum()), 1.0),
                         "dangerous_builds_per_step": float(cnt["dangerous"].sum()) / max(float(cnt["steps"].sum()), 1.0),
                         "max_pair_force": float(cnt["max_pair_force"].max()),
                         "phase_share": {k[4:]: v / cyc_tot for k, v in cyc.items()}},
            "clocks": sampler.summary(),
            "failed_replicas": int(fails[0]),
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(w, args.cpu_seconds)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "per_thread")}
            line["cpu_baseline"]["when"] = ("sequential: after the GPU legs, on all host threads (run concurrently it would take the core that "
                                            "drives the end-to-end leg; the GPU legs use one host thread)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="box12_p3_x512", choices=sorted(WORKLOADS))
    ap.add_argument("--replicas", type=int, default=0, help="replicas per GPU (default: the workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_native(args, w)


if __name__ == "__main__":
    main()
