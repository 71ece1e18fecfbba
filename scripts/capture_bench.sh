#!/bin/bash
# Round-end evidence (run on the GPU box): the bench without a profiler, then the ncu launch list and ONE full capture of the
# timed launch of the same command.  usage: scripts/capture_bench.sh TAG   ->  gpurun_out/TAG_*
TAG=$1
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:esb_run_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -c 400 gpurun_out/${TAG}_plain.json
