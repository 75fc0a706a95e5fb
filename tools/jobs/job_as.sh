for i in 1 2; do
NIG_ROLLOUT_PAIR=1 python tools/ab_rollout.py 65536 | sed "s/^/pair=1 /"
NIG_ROLLOUT_PAIR=1 NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_pairahead.so python tools/ab_rollout.py 65536 | sed "s/^/pair=1 /"
done > gpurun_out/r2_pair_ahead_64k.txt 2>&1
