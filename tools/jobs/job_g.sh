for n in 131072 262144 524288; do for g in 1 2 0; do NIG_GRID_STEP=$g python tools/grid_step_ab.py $n 0; done; done > gpurun_out/r2_grid_step_ab3.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
