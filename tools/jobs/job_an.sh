python -m pytest tests/test_gpu_rollout_host.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest_an.log
python bench.py --steps 20 --warmup 3 --sections e2e,configs > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err
NIG_HOST_DIRECT_MAX_MB=0,48 python bench.py --steps 20 --warmup 3 --sections e2e > gpurun_out/r2_bench15b.json 2> gpurun_out/r2_bench15b.err
