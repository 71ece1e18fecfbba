cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for R in 16 64 148 592; do timeout 120 python scripts/perf12.py $R 6 main | cut -c1-200; done
timeout 200 python scripts/perf_actin.py 64 80 | cut -c1-90
