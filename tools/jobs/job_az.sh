python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "grid" 2>&1 | tail -4 > gpurun_out/r2_pytest_az.log
for i in 1 2; do
python tools/grid_ab.py 1 1048576 | grep slices
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_gridahead.so python tools/grid_ab.py 1 1048576 | grep slices | sed "s/^/ahead /"
done > gpurun_out/r2_grid_ahead.txt 2>&1
