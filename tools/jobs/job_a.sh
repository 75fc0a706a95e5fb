set -x
tools/pipe_probe_bin 2000 new > gpurun_out/r2_probe2.txt 2>&1
python tools/grid_profile.py 1 1048576 64 > /dev/null 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_grid_roll python tools/grid_profile.py 1 1048576 64 > gpurun_out/r2_ncu_grid_roll.log 2>&1
ncu --page source --csv -i gpurun_out/r2_prof_grid_roll.ncu-rep > gpurun_out/r2_grid_roll_source.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:step_ -s 3 -c 1 -f -o gpurun_out/r2_prof_grid_step python tools/step_profile.py 2 > gpurun_out/r2_ncu_grid_step.log 2>&1
ncu --page source --csv -i gpurun_out/r2_prof_grid_step.ncu-rep > gpurun_out/r2_grid_step_source.csv 2>/dev/null
ncu --page details -i gpurun_out/r2_prof_grid_step.ncu-rep | head -40 > gpurun_out/r2_grid_step_details.txt
