python tools/k_sweep.py 65536 > gpurun_out/r2_k_sweep.txt 2>&1
