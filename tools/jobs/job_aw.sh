tools/launch_probe_bin > gpurun_out/r2_launch_probe.txt 2>&1
for t in 0 1; do python tools/step64k.py 65536 $t; python tools/step64k.py 16384 $t; done >> gpurun_out/r2_launch_probe.txt 2>&1
