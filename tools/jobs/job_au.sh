python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
