for n in 131072 196608 262144 524288; do for pr in 0 1; do NIG_ROLLOUT_PAIR=$pr python tools/ab_rollout.py $n | grep slices | sed "s/^/pair=$pr /"; done; done > gpurun_out/r2_pair_threshold.txt 2>&1
