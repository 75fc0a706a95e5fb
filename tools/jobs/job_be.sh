python tools/step64k.py 65536 0 > gpurun_out/r2_step_statsmodes.txt 2>&1
for v in sm1 sm2; do NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_$v.so python tools/step64k.py 65536 0 | sed "s/^/$v /"; done >> gpurun_out/r2_step_statsmodes.txt 2>&1
