python tools/step64k.py 65536 0 > gpurun_out/r2_step_statsearly.txt 2>&1
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_early.so python tools/step64k.py 65536 0 | sed "s/^/early /" >> gpurun_out/r2_step_statsearly.txt 2>&1
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_early.so python tools/step64k.py 65536 1 | sed "s/^/early /" >> gpurun_out/r2_step_statsearly.txt 2>&1
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_early.so python tools/step64k.py 16384 0 | sed "s/^/early /" >> gpurun_out/r2_step_statsearly.txt 2>&1
