python tools/host_direct_check.py > gpurun_out/r2_host_direct_1m.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.txt 2>&1
