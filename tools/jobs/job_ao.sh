for z in 1024 100000; do echo "NIG_ZERO_COPY_MAX_ENVS=$z"; NIG_ZERO_COPY_MAX_ENVS=$z python - <<'P'
import sys, os, json
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "neorl-industrial-gym_b200")]
import bench, types
import neorl_industrial as ni
args = types.SimpleNamespace(seed=0, steps=50, warmup=3)
for n in (4096, 16384, 65536):
    r = bench.e2e_step_api(ni, n, 0, args)
    print(n, r["value"], r["ms_per_call"], flush=True)
P
done > gpurun_out/r2_zero_copy_step.txt 2>&1
