python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "wrapper or bands or fast_loop" 2>&1 | tail -8 > gpurun_out/r2_pytest_u.log
python tools/wrapper_cost.py > gpurun_out/r2_wrapper_cost.txt 2>&1
