"""On-disk outputs of the reference drivers, written from the device's frame records.

Formats follow ``run_single_shoebox.py`` (== ``VisitorGeneMaps/run_single_GENES_STAGES.py``):
``run<r>/geneTrack.txt`` (:1496), ``run<r>/ser5pAroundCluster.txt`` (:743-748, 1578-1590), the event
log (:1595-1617), ``parallel_counter/run<r>.txt`` (:1619-1621), ``active_duration.txt`` (:1625-1627),
``gene_stats.txt`` (:1629-1642), ``{dist,ddist}_{induced,approaching,active,receding}.txt``
(:1644-1664) and the completion sentinel ``run<r>/figures/ser5p_cluster.pdf`` that the launcher polls
(``run_parallel_shoebox.sh:8-11,69-75``).  All numbers are printed with ``str()`` as the reference does.
"""
from __future__ import annotations

import os

import numpy as np

from . import protocol as P

EVENT_TYPES = ["induced", "approaching", "active", "receding"]
W1 = W2 = 20


def make_run_folders(out_folder: str, run_number: int, subfolders=("figures", "image_files", "microscopy_files")) -> str:
    """run_single_shoebox.py:308-317."""
    os.makedirs(os.path.join(out_folder, "parallel_counter"), exist_ok=True)
    run_dir = os.path.join(out_folder, f"run{run_number}")
    os.makedirs(run_dir, exist_ok=True)
    for s in subfolders:
        os.makedirs(os.path.join(run_dir, s), exist_ok=True)
    return run_dir


def gene_track_lines(frames: np.ndarray):
    """frames [n_chunks, 9] -> the lines of geneTrack.txt (:1496)."""
    for f in frames:
        yield (str(float(f[0])) + "," + str(float(f[1])) + "," + str(float(f[2])) + "," + str(float(f[3])) + "," + str(float(f[4])) + ","
               + str(float(f[5])) + "," + str(int(f[6])) + "," + str(int(f[7])) + "," + str(int(f[8])) + "\n")


def cluster_header() -> str:
    """:743-748."""
    r = P.CLUSTER_R[1:]
    return "t," + ",".join(str(v) for v in r) + "\n"


def event_log(states: np.ndarray) -> list:
    """:1595-1617.  0 induced but non-engaging, 1 approaching (w1 before), 2 active, 3 receding (w2 after)."""
    n = len(states)
    active_runs = {i for i in range(n) if int(states[i]) == 2}
    ev = [0] * n
    for i in range(n - 1):
        if i in active_runs:
            ev[i + 1] = 2
    for i in range(n):
        if (i - W2) >= 0:
            if ev[i - W2] == 2:
                if ev[i] != 2:
                    ev[i] = 3
        if (i + W1) <= (n - 1):
            if ev[i + W1] == 2:
                if ev[i] != 2:
                    ev[i] = 1
    return ev


def actin_track_lines(stats: np.ndarray):
    """actinTrack.txt rows (run_single_shoebox.py:1544): t, n_filaments, round(mean length, 2), round(sd, 2), n_free,
    tab separated.  stats [n, 6] = n_filaments, sum len, sum len^2, n_free, min, max; mean/sd are np.mean / np.std of
    the integer lengths, recomputed here from the two sums (equal to the reference's values after the rounding to 2
    decimals except when the third decimal sits on a rounding boundary)."""
    for i, (nf, s1, s2, nfree, _, _) in enumerate(np.asarray(stats, np.int64)):
        if nf > 0:
            mean = float(s1) / float(nf)
            sd = float(np.sqrt(max(float(s2) / float(nf) - mean * mean, 0.0)))
        else:
            mean = sd = float("nan")
        yield str((i + 1) * P.DELT) + "\t" + str(int(nf)) + "\t" + str(round(mean, 2)) + "\t" + str(round(sd, 2)) + "\t" + str(int(nfree)) + "\n"


def write_actin(out_folder: str, run_number: int, stats: np.ndarray, hist: np.ndarray) -> None:
    """actinTrack.txt and actinAroundCluster.txt of one repeat (run_single_shoebox.py:738-756, 1544, 1581-1590)."""
    run_dir = os.path.join(out_folder, "run" + str(run_number))
    with open(os.path.join(run_dir, "actinTrack.txt"), "w") as f:
        f.writelines(actin_track_lines(stats))
    with open(os.path.join(run_dir, "actinAroundCluster.txt"), "w") as f:
        f.write(cluster_header())
        for i in range(len(stats)):
            f.write(str((i + 1) * P.DELT) + "," + ",".join(str(int(v)) for v in hist[i]) + "\n")


def microscopy_stack(counts: np.ndarray, sigma: float = P.MICROSCOPY_SIGMA) -> np.ndarray:
    """(C, Z, Y, X) float32 image stack of one frame from its voxel counts [5, nx, ny, nz]:
    gaussian_filter(np.transpose(H, (2, 1, 0)), sigma).astype(np.float32) per channel
    (run_single_GENE_STAGES_microscopyModif.py:326-353)."""
    from scipy.ndimage import gaussian_filter
    return np.stack([gaussian_filter(np.transpose(np.asarray(h, np.float64), (2, 1, 0)), sigma=sigma).astype(np.float32) for h in counts], axis=0)


def write_microscopy(out_folder: str, run_number: int, counts: np.ndarray, chunk_indices, t_run: int = P.T_RUN) -> None:
    """microscopy_files/synthetic_microscopy_<time>.npy (and .ome.tif when tifffile is installed) per recorded chunk:
    the array the reference hands to tifffile.imwrite / imshow.  time = (i+1)*tRun zero-filled to 7 digits
    (run_single_shoebox.py:1502).  Rendering PNGs with matplotlib is left to the reference's plotting scripts."""
    d = os.path.join(out_folder, "run" + str(run_number), "microscopy_files")
    os.makedirs(d, exist_ok=True)
    try:
        import tifffile
    except Exception:
        tifffile = None
    for h, i in zip(counts, chunk_indices):
        stack = microscopy_stack(h)
        name = os.path.join(d, "synthetic_microscopy_" + str((i + 1) * t_run).zfill(7))
        np.save(name + ".npy", stack)
        if tifffile is not None:
            tifffile.imwrite(name + ".ome.tif", stack, photometric="minisblack",
                             metadata={"axes": "CZYX", "Channel": {"Name": [c[0] for c in P.MICROSCOPY_CHANNELS]}})


def write_run(out_folder: str, run_number: int, frames: np.ndarray, hist: np.ndarray, raw: np.ndarray, total_active_steps: int,
              n_runs: int = P.N_RUNS, subfolders=("figures", "image_files", "microscopy_files")) -> None:
    """Everything one repeat leaves on disk.  frames [n,9], hist [n,49], raw [n,2] = (d_rp, d_rg) in bead units."""
    run_dir = make_run_folders(out_folder, run_number, subfolders)
    with open(os.path.join(run_dir, "geneTrack.txt"), "w") as f:
        f.writelines(gene_track_lines(frames))
    with open(os.path.join(run_dir, "ser5pAroundCluster.txt"), "w") as f:
        f.write(cluster_header())
        for i in range(len(frames)):
            f.write(str((i + 1) * P.DELT) + "," + ",".join(str(int(v)) for v in hist[i]) + "\n")
    ev = event_log(frames[:, 8])
    with open(os.path.join(out_folder, "parallel_counter", f"run{run_number}.txt"), "w") as f:
        f.write("Run " + str(run_number) + " done!\n")
    prog = os.path.join(out_folder, "parallel_counter", f"progress_run{run_number}.txt")
    if os.path.exists(prog):
        os.remove(prog)
    with open(os.path.join(out_folder, "active_duration.txt"), "a") as f:
        f.write(f"{run_number},{total_active_steps},{n_runs - P.T_INDUCTION_ON}\n")
    with open(os.path.join(out_folder, "gene_stats.txt"), "a") as f:
        for j in range(len(frames)):
            if j < P.T_INDUCTION_ON:
                continue
            d = float(raw[j, 1])
            in_contact = 1 if d < P.POL2_RELEASE_RADIUS else 0
            s2p = 1 if ev[j] == 2 else 0
            f.write(str(in_contact) + "," + str(d * P.SIG_CHROMATIN_NM) + "," + str(int(frames[j, 7])) + "," + str(s2p) + "\n")
    dist = [float(raw[x, 1]) * P.SIG_CHROMATIN_NM for x in range(len(frames))]
    ddist = [dist[k + 1] - dist[k] for k in range(len(dist) - 1)]
    for name, arr in (("dist", dist), ("ddist", ddist)):
        for j, ev_name in enumerate(EVENT_TYPES):
            with open(os.path.join(out_folder, f"{name}_{ev_name}.txt"), "a") as f:
                for i in range(len(arr)):
                    if ev[i] == j:
                        f.write(str(arr[i]) + "\n")
    # completion sentinel the launcher waits for (it deletes the file afterwards)
    with open(os.path.join(run_dir, "figures", "ser5p_cluster.pdf"), "wb") as f:
        f.write(b"%PDF-1.1\n% placeholder written by shoebox-b200: plots are out of scope, the launcher only needs the file\n%%EOF\n")
