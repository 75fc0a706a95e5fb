python tools/ab_rollout.py 65536 | grep slices > gpurun_out/r2_rollout_live.txt 2>&1
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_live.so python tools/ab_rollout.py 65536 | grep slices | sed "s/^/counters-live-no-flush /" >> gpurun_out/r2_rollout_live.txt 2>&1
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_norollstats.so python tools/ab_rollout.py 65536 | grep slices | sed "s/^/nostats /" >> gpurun_out/r2_rollout_live.txt 2>&1
