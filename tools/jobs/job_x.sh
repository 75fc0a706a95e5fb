for g in 1 0; do NIG_STEPS_GRAPH=$g python tools/ab_rollout.py 65536 | sed "s/^/steps_graph=$g /"; done > gpurun_out/r2_steps_graph_ab.txt 2>&1
for k in 5 20; do for g in 1 0; do echo "steps=$k graph=$g"; NIG_STEPS_GRAPH=$g python bench.py --steps $k --warmup 3 --sections none 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['gpu_launches'])
"; done; done >> gpurun_out/r2_steps_graph_ab.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_x.log
