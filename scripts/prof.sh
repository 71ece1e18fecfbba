#!/bin/bash
# ncu capture of the timed launch of scripts/perf12.py (run on the GPU box): usage scripts/prof.sh TAG [R] [chunks]
TAG=$1; R=${2:-512}; NCH=${3:-2}
mkdir -p gpurun_out
python scripts/perf12.py $R $NCH $TAG > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:esb_run_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/${TAG}_prof \
    python scripts/perf12.py $R $NCH $TAG > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_plain.log
