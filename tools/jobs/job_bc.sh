for i in 1 2; do
python tools/step64k.py 65536 0
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_nostats.so python tools/step64k.py 65536 0 | sed "s/^/nostats /"
done > gpurun_out/r2_step_nostats.txt 2>&1
