for i in 1 2; do
python tools/ab_rollout.py 65536 | grep slices
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_norollstats.so python tools/ab_rollout.py 65536 | grep slices | sed "s/^/nostats /"
done > gpurun_out/r2_rollout_nostats.txt 2>&1
