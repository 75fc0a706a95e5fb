python -m pytest tests/test_gpu_round2.py -x -q -k "grid" 2>&1 | tail -15 > gpurun_out/r2_pytest_grid.log
python -m pytest tests/test_gpu_parity.py -x -q -k "grid" 2>&1 | tail -5 >> gpurun_out/r2_pytest_grid.log
for g in 1 2; do for tr in 0 1; do NIG_GRID_STEP=$g python tools/grid_step_ab.py 1048576 $tr; done; done > gpurun_out/r2_grid_step_ab2.txt 2>&1
for g in 1 2; do NIG_GRID_STEP=$g python tools/grid_step_ab.py 4194304 0; done >> gpurun_out/r2_grid_step_ab2.txt 2>&1
NIG_GRID_FAST=1 python tools/grid_ab.py 1 1048576 >> gpurun_out/r2_grid_step_ab2.txt 2>&1
