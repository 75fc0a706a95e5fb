python -m pytest tests/test_gpu_round2.py -x -q -k "grid" 2>&1 | tail -5 > gpurun_out/r2_pytest_grid.log
for s in 3 4 5 6 7 8 9; do NIG_GRID_FAST=$s python tools/grid_ab.py 1 1048576; done > gpurun_out/r2_grid_ab3.txt 2>&1
