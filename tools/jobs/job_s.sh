for i in 1 2; do
python tools/ab_rollout.py 65536 1048576
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_ahead.so python tools/ab_rollout.py 65536 1048576
done > gpurun_out/r2_ahead_ab.txt 2>&1
