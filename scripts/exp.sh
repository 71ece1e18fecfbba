cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python scripts/perf12.py 592 6 after_actin
