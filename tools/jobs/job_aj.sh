ncu --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 12 --launch-count 1 -f -o gpurun_out/r2h_prof_reactor_w1 python tools/prof_rollout.py 65536 32 > gpurun_out/r2h_ncu_reactor_w1.log 2>&1
ncu --page source --csv -i gpurun_out/r2h_prof_reactor_w1.ncu-rep > gpurun_out/r2h_reactor_w1_source.csv 2>/dev/null
