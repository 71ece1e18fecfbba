cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ESB_WIDE=0 timeout 200 python scripts/perf_actin.py 64 80
ESB_WIDE=1 timeout 200 python scripts/perf_actin.py 64 80
timeout 200 python scripts/perf_actin.py 148 80
timeout 120 python scripts/perf12.py 592 6 v512
