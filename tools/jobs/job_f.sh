python -m pytest tests/test_gpu_round2.py -x -q -k "grid_persistent" 2>&1 | tail -15 > gpurun_out/r2_pytest_grid.log
python -m pytest tests/test_gpu_parity.py -x -q -k "grid" 2>&1 | tail -5 >> gpurun_out/r2_pytest_grid.log
for tr in 0 1; do NIG_GRID_STEP=1 python tools/grid_step_ab.py 1048576 $tr; done > gpurun_out/r2_grid_step_ab2.txt 2>&1
NIG_GRID_STEP=1 python tools/grid_step_ab.py 4194304 0 >> gpurun_out/r2_grid_step_ab2.txt 2>&1
NIG_GRID_STEP=1 python tools/grid_step_ab.py 262144 0 >> gpurun_out/r2_grid_step_ab2.txt 2>&1
NIG_GRID_STEP=0 python tools/grid_step_ab.py 262144 0 >> gpurun_out/r2_grid_step_ab2.txt 2>&1
