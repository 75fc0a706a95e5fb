python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_pytest_bi.log
