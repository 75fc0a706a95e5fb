python tools/prof_rollout.py 65536 > /dev/null 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 12 --launch-count 1 -f -o gpurun_out/r2h_prof_reactor python tools/prof_rollout.py 65536 > gpurun_out/r2h_ncu_reactor.log 2>&1
ncu --page source --csv -i gpurun_out/r2h_prof_reactor.ncu-rep > gpurun_out/r2h_reactor_source.csv 2>/dev/null
python bench.py --steps 2 --warmup 3 --sections none > gpurun_out/r2h_b.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --sections none > gpurun_out/r2h_ncu_b.log 2>&1
