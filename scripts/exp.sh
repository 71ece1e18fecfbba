cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_s2c.log 2>&1; tail -1 gpurun_out/bench_s2c.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:esb_run_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r1c_bench python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
