python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h_pytest_multi.log
