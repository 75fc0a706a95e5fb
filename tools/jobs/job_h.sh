python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_m.log
ncu --set full --import-source on --clock-control none -k regex:rollout_grid_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_grid_fast python tools/grid_profile.py 1 1048576 64 > gpurun_out/r2_ncu_grid_fast.log 2>&1
ncu --page source --csv -i gpurun_out/r2_prof_grid_fast.ncu-rep > gpurun_out/r2_grid_fast_source.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:step_grid_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_grid_step python tools/step_profile.py 2 > gpurun_out/r2_ncu_grid_step.log 2>&1
ncu --page source --csv -i gpurun_out/r2_prof_grid_step.ncu-rep > gpurun_out/r2_grid_step_source.csv 2>/dev/null
