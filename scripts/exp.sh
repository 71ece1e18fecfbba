cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K=10 value %.4e e2e %.4e' % (d['value'], d['e2e']['value']))"
timeout 120 python scripts/perf12.py 592 6 seg_tmpl
