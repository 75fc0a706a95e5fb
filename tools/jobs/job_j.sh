python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
python tools/host_path_breakdown.py > gpurun_out/r2_hostpath2.txt 2>&1
