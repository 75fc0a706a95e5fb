python tools/host_path_breakdown.py > gpurun_out/r2_hostpath4.txt 2>&1
