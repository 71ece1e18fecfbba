cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python scripts/perf12.py 592 6 main
timeout 200 python scripts/perf_actin.py 64 80 | cut -c1-90
