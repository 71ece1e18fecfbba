cd /root/repo
timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q 2>&1 | tail -8
timeout 600 python run_sweep.py -b 12 --promoters 3 --activations 30 --thresholds 10 20 30 40 50 60 70 80 90 100 --repeats 55 -t 400 -o gpurun_out/sweep_box12.txt 2>&1 | tail -5
head -5 gpurun_out/sweep_box12.txt
