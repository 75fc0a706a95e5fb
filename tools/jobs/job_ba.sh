python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.txt 2>&1
python bench.py --steps 20 --warmup 3 --sections configs > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err
