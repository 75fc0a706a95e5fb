python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_v.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
