NIG_MIN_SLICE=4096 python tools/host_direct_sweep.py > gpurun_out/r2_host_direct_sweep.txt 2>&1
