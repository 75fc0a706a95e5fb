for s in 8 10 12 16; do NIG_MIN_SLICE=2048 NIG_HOST_SLICES=$s python tools/ab_rollout.py 65536 | grep slices | sed "s/^/host_slices=$s /"; done > gpurun_out/r2_slices_graph.txt 2>&1
