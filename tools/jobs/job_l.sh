python -m pytest tests/test_gpu_round2.py -x -q -k "grid" 2>&1 | tail -5 > gpurun_out/r2_pytest_grid.log
for g in 1 3 4 5 2; do NIG_GRID_STEP=$g python tools/grid_step_ab.py 1048576 0; done > gpurun_out/r2_grid_step_ab4.txt 2>&1
for g in 1 3 5; do NIG_GRID_STEP=$g python tools/grid_step_ab.py 4194304 0; done >> gpurun_out/r2_grid_step_ab4.txt 2>&1
for g in 1 3 5; do NIG_GRID_STEP=$g python tools/grid_step_ab.py 1048576 1; done >> gpurun_out/r2_grid_step_ab4.txt 2>&1
