import os, sys
sys.argv = ["x"]
exec(open("tools/perf_sweep.py").read().split("for kind, name in")[0])
for n in (1 << 20, 1 << 22, 1 << 24):
    for vec in (1, 2):
        gbs, ms = run(0, n, vec)
        print(f"reactor n={n:9d} vec={vec}: {gbs:7.0f} GB/s ({ms*1e3:8.1f} us) frac {gbs/6450.9:.3f}", flush=True)
