#!/usr/bin/env python
"""The GENE_STAGES sweep on the GPUs of one box: what VisitorGeneMaps/run_singleCond_parallel.sh (one process per
repeat) followed by dist_genestages_grouped.py does through the filesystem, as one job.

    python run_sweep.py -b 12 --promoters 1 2 3 --activations 30 --thresholds 10 20 ... --repeats 100 -o summary.txt
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 run_sweep.py ...       (one rank per GPU)

Writes summary_contact_grouped_*.txt in the reference's format (rank 0)."""
import argparse
import os
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-b", "--box", type=int, nargs="+", default=[12], help="box size(s); with several, one table per box: <out>.box<B>")
    ap.add_argument("--promoters", type=int, nargs="+", default=[1, 2, 3])
    ap.add_argument("--activations", type=int, nargs="+", default=[1, 5, 10, 15, 20, 25, 30, 50, 75, 100])
    ap.add_argument("--thresholds", type=int, nargs="+", default=[10, 20, 30, 40, 50, 60, 70, 75, 80, 90, 100])
    ap.add_argument("--repeats", type=int, default=100)
    ap.add_argument("-t", "--total", type=int, default=None, help="Python timesteps per run (default 2000)")
    ap.add_argument("--batch-groups", type=int, default=118, help="5-run groups per batch (118 x 5 = 590 replicas = two CTAs per SM)")
    ap.add_argument("-o", "--out", default="summary_contact_grouped_Thresholds10-100.txt")
    ap.add_argument("--cpp", action="store_true", help="write the table in the format of dist_genestages_grouped_cpp (no index column, 6 significant digits)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from enhancershoebox_b200 import sweep
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for box in args.box:
        t0 = time.time()
        groups, rows = sweep.run_sweep(box, args.promoters, args.activations, args.thresholds, args.repeats, args.total,
                                       batch_groups=args.batch_groups, device=local,
                                       progress=lambda r, p, n, left: print(f"rank {r}: box {box} promoter {p}: {n} groups done, {left} queued", flush=True))
        if (dist.get_rank() if dist.is_initialized() else 0) == 0:
            out = args.out if len(args.box) == 1 else f"{args.out}.box{box}"
            with open(out, "w") as f:
                f.write(sweep.summary_cpp_csv(groups, rows) if args.cpp else sweep.summary_csv(groups, rows))
            print(f"box {box}: {len(groups)} rows -> {out} in {time.time() - t0:.1f} s")
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
