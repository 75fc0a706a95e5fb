python tools/ab_rollout.py 65536 | grep slices > gpurun_out/r2_loop_probe.txt 2>&1
for v in l1 l2; do NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_$v.so python tools/ab_rollout.py 65536 | grep slices | sed "s/^/$v /"; done >> gpurun_out/r2_loop_probe.txt 2>&1
