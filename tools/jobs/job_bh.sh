python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "without_device_counters" 2>&1 | tail -8 > gpurun_out/r2_pytest_bh.log
for c in 1 0; do python tools/step64k.py 65536 0 $c; python tools/step64k.py 16384 0 $c; done > gpurun_out/r2_step_counters.txt 2>&1
python bench.py --steps 5 --warmup 3 --sections single 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)['single_step']; print(json.dumps({k: d[k] for k in d if k.startswith('envs_64k')}, indent=1))" >> gpurun_out/r2_step_counters.txt 2>&1
