python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_ai.log
python bench.py --steps 20 --warmup 3 --sections none > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err
