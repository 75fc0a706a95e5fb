python tools/host_direct_check.py > gpurun_out/r2_host_direct_sizes.txt 2>&1
