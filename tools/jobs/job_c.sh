python -m pytest tests/test_gpu_round2.py -x -q -k "grid" 2>&1 | tail -15 > gpurun_out/r2_pytest_grid.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout_host.py -x -q -k "grid" 2>&1 | tail -15 >> gpurun_out/r2_pytest_grid.log
for s in 4 5 6 7 8; do NIG_GRID_FAST=$s python tools/grid_ab.py 1 1048576; done > gpurun_out/r2_grid_ab2.txt 2>&1
NIG_GRID_FAST=4 ncu --set full --import-source on --clock-control none -k regex:rollout_grid_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_grid_fast python tools/grid_profile.py 1 1048576 64 > gpurun_out/r2_ncu_grid_fast.log 2>&1
ncu --page source --csv -i gpurun_out/r2_prof_grid_fast.ncu-rep > gpurun_out/r2_grid_fast_source.csv 2>/dev/null
