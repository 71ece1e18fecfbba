cd /root/repo
timeout 300 python -m pytest tests/test_gpu_forces.py tests/test_gpu_trajectory.py tests/test_gpu_ensemble.py -x -q 2>&1 | tail -3
for l in lds pair2; do ESB_CELL=2.8 ESB_LIB=variants/libesb_$l.so timeout 120 python scripts/perf12.py 592 6 ${l}_c2.8; done
