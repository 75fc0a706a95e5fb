for i in 1 2; do
python tools/ab_rollout.py 65536
for v in onevote unroll2 ov_u2; do
NIG_LIB_PATH=$PWD/neorl-industrial-gym_b200/_ab/libnig_b200_$v.so python tools/ab_rollout.py 65536
done
done > gpurun_out/r2_ahead_ab2.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_t.log
