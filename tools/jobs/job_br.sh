python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --sections e2e > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2k_pytest_multi.log
