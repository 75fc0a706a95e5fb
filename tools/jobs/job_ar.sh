python tools/soak_parity.py 500003 12 > gpurun_out/r2h_soak_parity.txt 2>&1
