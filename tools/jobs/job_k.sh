for pdl in 0 1 0 1; do echo "== NIG_ROLLOUT_PDL=$pdl"; NIG_ROLLOUT_PDL=$pdl python tools/host_path_breakdown.py; done > gpurun_out/r2_hostpath3.txt 2>&1
