python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_y.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
