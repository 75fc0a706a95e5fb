for i in 1 2; do
python tools/step64k.py 65536 0
for sh in 512 2048; do NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_sh$sh.so python tools/step64k.py 65536 0 | sed "s/^/shards=$sh /"; done
done > gpurun_out/r2_step_shards.txt 2>&1
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_sh2048.so python tools/step64k.py 65536 1 | sed "s/^/shards=2048 /" >> gpurun_out/r2_step_shards.txt 2>&1
