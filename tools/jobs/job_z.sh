python -m pytest tests/test_gpu_rollout_host.py tests/test_gpu_reference_suite.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_z.log
for d in 1 0 1 0; do echo "== NIG_HOST_DIRECT=$d"; NIG_HOST_DIRECT=$d python tools/host_path_breakdown.py 2>&1 | grep "graph=1"; done > gpurun_out/r2_hostpath5.txt 2>&1
