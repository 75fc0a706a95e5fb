tools/launch_probe_bin > gpurun_out/r2_launch_probe2.txt 2>&1
