python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_aa.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err
