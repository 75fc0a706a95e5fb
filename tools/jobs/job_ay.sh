for i in 1 2; do
python tools/grid_ab.py 1 1048576
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_lanediv.so python tools/grid_ab.py 1 1048576 | sed "s/^/lanediv /"
done > gpurun_out/r2_grid_lanediv.txt 2>&1
