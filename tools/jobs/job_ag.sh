python -m pytest tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_ag.log
python bench.py --steps 5 --warmup 3 --sections configs > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err
