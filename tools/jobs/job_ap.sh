python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_ap.log
python tools/single_env_latency.py > gpurun_out/r2_single_env_latency.txt 2>&1
