python tools/ab_rollout.py 65536 > gpurun_out/r2_direct_stats.txt 2>&1
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_rollout_host.py -m gpu -x -q -k "reactor or fast_loop or bands or supplied or graph_replay or direct" 2>&1 | tail -5 >> gpurun_out/r2_direct_stats.txt
