python tools/ab_rollout.py 65536 > gpurun_out/r2_ab_w.txt 2>&1
python bench.py > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err
python tools/ab_rollout.py 65536 >> gpurun_out/r2_ab_w.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_w.log
