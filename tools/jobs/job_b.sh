python -m pytest tests/test_gpu_round2.py -x -q -k "grid" 2>&1 | tail -15 > gpurun_out/r2_pytest_grid.log
python -m pytest tests/test_gpu_parity.py -x -q -k "grid" 2>&1 | tail -15 >> gpurun_out/r2_pytest_grid.log
for s in 0 1 2 3 4; do NIG_GRID_FAST=$s python tools/grid_ab.py 1 1048576 131072; done > gpurun_out/r2_grid_ab.txt 2>&1
