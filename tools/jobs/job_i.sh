for pdl in 1 0 1 0; do NIG_ROLLOUT_PDL=$pdl python tools/ab_rollout.py 65536 1048576 | sed "s/^/pdl=$pdl /"; done > gpurun_out/r2_pdl_ab.txt 2>&1
NIG_ROLLOUT_PDL=1 python tools/grid_ab.py 1 1048576 >> gpurun_out/r2_pdl_ab.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_n.log
