"""nig_rollout_steps at 65,536 reactor envs: env slices x CTA size (NIG_HOST_SLICES / NIG_MIN_SLICE / NIG_ROLLOUT_BLOCK)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for slices in (1, 2, 4, 8, 16):
    for block in (32, 64, 128):
        env = dict(os.environ, NIG_HOST_SLICES=str(slices), NIG_MIN_SLICE="4096", NIG_ROLLOUT_BLOCK=str(block), NIG_ROLLOUT_PAIR="0")
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_rollout.py"), "65536"], env=env, capture_output=True, text=True).stdout
        line = [l for l in out.splitlines() if "slices" in l]
        print(f"slices={slices:2d} block={block:3d}: {line[0].split('median')[1].strip() if line else out[-200:]}", flush=True)
