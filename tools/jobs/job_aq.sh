python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "graph_replay" 2>&1 | tail -12 > gpurun_out/r2_pytest_aq.log
for qq in 1 0 1 0; do NIG_ROLLOUT_QUEUE=$qq python tools/ab_rollout.py 65536 | grep slices | sed "s/^/queue=$qq /"; done > gpurun_out/r2_queue_ab.txt 2>&1
