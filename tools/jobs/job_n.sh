ncu --set full --import-source on --clock-control none -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_robot_roll python tools/grid_profile.py 2 1048576 64 > gpurun_out/r2_ncu_robot_roll.log 2>&1
ncu --page source --csv -i gpurun_out/r2_prof_robot_roll.ncu-rep > gpurun_out/r2_robot_roll_source.csv 2>/dev/null
python tools/grid_ab.py 2 1048576 > gpurun_out/r2_robot_ab.txt 2>&1
