python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench8_ref.json 2> gpurun_out/r2_bench8_ref.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest_o.log
