for sh in 1 0 1 0; do NIG_ROLLOUT_SHARDS=$sh python tools/ab_rollout.py 65536 | sed "s/^/shards=$sh /"; done > gpurun_out/r2_shards_ab.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_p.log
