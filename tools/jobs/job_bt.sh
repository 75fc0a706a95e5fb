python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.txt 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_final3.log
python bench.py --steps 20 --warmup 3 --sections none > gpurun_out/r2_bench_final_headline.json 2>/dev/null
