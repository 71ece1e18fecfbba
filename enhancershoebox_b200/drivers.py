"""Drop-in command lines of the reference's two per-repeat drivers, backed by the GPU batch engine.

    python run_single_shoebox.py      -b BOX -r REPEAT -t TOTAL -o OUT/ -m 0|1 -c CONDITION -p PROMOTER -a ACT -x THR [-z NACTIN -n PNUC]
    python run_single_GENES_STAGES.py -b BOX -r REPEAT -t TOTAL -o OUT/ -m 0|1 -c CONDITION -p PROMOTER -a ACT -x THR

Same flags as ``run_single_shoebox.py:236-300`` / ``VisitorGeneMaps/run_single_GENES_STAGES.py:316-342``
(``-a`` is divided by 1000, :282), same inputs from the working directory
(``input_files/IC_...data``, :449-453; ``./LJsoft_RSQ.table``, :803 -- regenerated in memory when the
57 MB file is absent), same outputs (see writers.py).  One extension: ``-r A:B`` runs repeats A..B
(inclusive) as ONE batch on the GPU, which is how the engine is meant to be used; a single ``-r R``
works but occupies one CTA.  Repeats of a condition share the data file and differ in the seed, as
in the reference (which draws ``random.randint(1, 100000)`` per process, :642-644; here the seed is
``ESB_SEED`` (default 0) * 1_000_003 + repeat so that runs are reproducible).

Failure semantics: LAMMPS aborts a run that loses an atom / hits a wall and the launcher deletes the
run folder and starts the repeat again (``run_parallel_shoebox.sh:1-14``).  Here a failed replica is
re-run in-process with a fresh seed (up to ESB_RETRIES, default 5); if it still fails nothing is
written for that repeat and the exit status is 1, so the launcher's rule applies unchanged.
"""
from __future__ import annotations

import getopt
import os
import sys

import numpy as np

from . import datafile, protocol as P, writers
from .engine import ShoeboxBatch

USAGE = ("-b <box_size> -r <repeat | first:last> -t <total_runs> -o <out_folder> -m <make_images> -c <condition> "
         "-p <promoter_length> -a <activation_rate> -x <threshold>")


def parse_args(argv, with_actin: bool):
    short = "hb:r:t:o:m:c:p:a:x:" + ("z:n:" if with_actin else "")
    long = ["box=", "repeat=", "total=", "outfolder=", "make_images=", "condition=", "promoter=", "activation=", "threshold="]
    if with_actin:
        long += ["actin=", "nucleation="]
    try:
        opts, _ = getopt.getopt(argv, short, long)
    except getopt.GetoptError:
        print(USAGE)
        sys.exit(2)
    a = dict(box=11, repeats=[1], total=1, out="out/", make_images=0, condition="Control", promoter=3, activation=30.0, threshold=80,
             actin=0, p_nucleation=0.0)
    for opt, arg in opts:
        if opt == "-h":
            print(USAGE)
            sys.exit()
        elif opt in ("-b", "--box"):
            a["box"] = int(arg)
        elif opt in ("-r", "--repeat"):
            if ":" in arg:
                lo, hi = arg.split(":")
                a["repeats"] = list(range(int(lo), int(hi) + 1))
            else:
                a["repeats"] = [int(arg)]
        elif opt in ("-t", "--total"):
            a["total"] = int(arg)
        elif opt in ("-o", "--outfolder"):
            a["out"] = arg
        elif opt in ("-m", "--make_images"):
            a["make_images"] = int(arg)
        elif opt in ("-c", "--condition"):
            a["condition"] = arg
        elif opt in ("-p", "--promoter"):
            a["promoter"] = int(arg)
        elif opt in ("-a", "--activation"):
            a["activation"] = float(arg)
        elif opt in ("-x", "--threshold"):
            a["threshold"] = int(arg)
        elif opt in ("-z", "--actin"):
            a["actin"] = int(arg)
        elif opt in ("-n", "--nucleation"):
            a["p_nucleation"] = float(arg)
    return a


def run_condition(driver: str, a: dict, n_chunks: int | None = None, device: int = 0, data=None, quiet: bool = False,
                  make_microscopy: bool = False) -> int:
    """Run the repeats of one condition as a batch and write their folders.  Returns the number of
    repeats that could not be completed."""
    condition = a["condition"]
    is_actin = driver == "shoebox" and a.get("actin", 0) > 0 and condition != "noactin"
    # actin polymerisation events (run_single_shoebox.py:969-1298) run in every actin condition except the
    # static-topology ones, LatB and SwinA_nopol (:459-460)
    p_nucleation = float(a.get("p_nucleation", 0.0)) if is_actin else None
    if data is None:
        path = datafile.input_file_name(a["box"], a["promoter"], is_actin, a.get("actin") or None)
        if not os.path.exists(path):
            raise SystemExit(f"{path} not found: create it with make_input_file.py (as for the reference)")
        data = datafile.read_data(path)
    table_file = "LJsoft_RSQ.table" if os.path.exists("LJsoft_RSQ.table") else None
    total = (P.N_RUNS + P.TREATMENT_DURATION.get(condition, 0)) if n_chunks is None else n_chunks
    repeats = list(a["repeats"])
    base = int(os.environ.get("ESB_SEED", "0")) * 1_000_003
    retries = int(os.environ.get("ESB_RETRIES", "5"))
    subfolders = ("figures", "image_files", "microscopy_files") if driver == "shoebox" else ("figures", "image_files", "summaries")
    pending, attempt, failed = repeats, 0, []
    while pending:
        seeds = np.array([base + r + 7919 * attempt * 100_003 for r in pending], dtype=np.uint64)
        for r in pending:
            writers.make_run_folders(a["out"], r, subfolders)
            with open(os.path.join(a["out"], "parallel_counter", f"progress_run{r}.txt"), "w") as f:
                f.write(f"{r},0")
        eng = ShoeboxBatch(data, n_replicas=len(pending), driver=driver, condition=condition, device=device, table_file=table_file,
                           seeds=seeds, thresholds=a["threshold"], activations=a["activation"], p_nucleation=p_nucleation)
        eng.set_schedule(total, observe=True, microscopy=make_microscopy)
        eng.run()
        status = eng.status()
        mic = eng.microscopy() if make_microscopy and eng.mic_chunks else None
        frames, hist = eng.frames(with_hist=True)
        actin = eng.actin_frames() if eng.actin is not None else None
        raw = eng.raw_distances()
        active = eng.counters()["total_active"]
        eng.close()
        again = []
        for k, r in enumerate(pending):
            if status[k] != 0:
                again.append(r)
                continue
            writers.write_run(a["out"], r, frames[k], hist[k], raw[k], int(active[k]), subfolders=subfolders)
            if actin is not None:
                writers.write_actin(a["out"], r, actin[0][k], actin[1][k])
            if mic is not None:
                writers.write_microscopy(a["out"], r, mic[k], eng.mic_chunks)
            if not quiet:
                print(f"Run {r} done ({int((frames[k][:, 8] == 2).sum())} active frames)")
        attempt += 1
        if again and attempt > retries:
            failed = again
            break
        pending = again
    return len(failed)


def pick_device(repeats, n_devices: int | None = None) -> int:
    """GPU of this process.  The reference's launchers start one OS process per repeat, 8 or 9 at a time
    (run_parallel_shoebox.sh:58-76, VisitorGeneMaps/run_singleCond_parallel.sh:38-61): under the UNCHANGED launcher the
    processes spread over the GPUs of the box by repeat number (repeat % n_gpus).  ESB_DEVICE pins a device;
    CUDA_VISIBLE_DEVICES narrows the set as usual."""
    if "ESB_DEVICE" in os.environ:
        return int(os.environ["ESB_DEVICE"])
    if n_devices is None:
        try:
            import ctypes
            cudart = ctypes.CDLL("libcudart.so")
            cnt = ctypes.c_int(0)
            n_devices = cnt.value if cudart.cudaGetDeviceCount(ctypes.byref(cnt)) == 0 and cnt.value > 0 else 1
        except OSError:
            try:
                import torch
                n_devices = max(torch.cuda.device_count(), 1)
            except Exception:
                n_devices = 1
    return int(repeats[0]) % max(n_devices, 1)


def main_shoebox(argv=None) -> int:
    a = parse_args(sys.argv[1:] if argv is None else argv, with_actin=True)
    return 1 if run_condition("shoebox", a, device=pick_device(a["repeats"])) else 0


def main_genes_stages(argv=None, make_microscopy: bool = False) -> int:
    a = parse_args(sys.argv[1:] if argv is None else argv, with_actin=False)
    return 1 if run_condition("genes_stages", a, make_microscopy=make_microscopy, device=pick_device(a["repeats"])) else 0
