python -m pytest tests/test_gpu_rollout_host.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_ah.log
