"""Worker of tests/test_gpu_multi.py: one rank per GPU under torchrun, sharded fused rollout + the NCCL all-reduce of the
device stats block; rank 0 compares with the unsharded run on its own GPU and prints OK."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np
import torch
import torch.distributed as dist

import neorl_industrial as ni
from neorl_industrial import _native as N
from neorl_industrial.distributed import allreduce_device_stats, allreduce_extrema, make_sharded, world_info

rank, world, local = world_info()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total, K = 100_003, 96
from neorl_industrial.safety import BoundConstraint, SafetyWrapper
env = SafetyWrapper(make_sharded("ChemicalReactor-v0", n_total, seed=5),
                    constraints=[BoundConstraint("temperature_band", 0, 280.0, 321.0, penalty=-100.0)])
nat = env.native
nat.reset_device()
nat.track_extrema(True)
nat.rollout_device(K, N.POLICY_UNIFORM)
view = allreduce_device_stats(nat)                     # in place on the device block, int64 counters + fp64 sums
torch.cuda.synchronize()
summed = nat.stats_dict()
nat.rollout_device(512, N.POLICY_UNIFORM)              # past the 500-step truncation: every env finishes an episode
ext = allreduce_extrema(nat)                           # one MAX all-reduce of the two int64 keys
state = nat.get_state_host()[0]
gathered = [None] * world
dist.all_gather_object(gathered, state)
if rank == 0:
    whole = SafetyWrapper(ni.make("ChemicalReactor-v0", num_envs=n_total, seed=5, device=f"cuda:{local}"),
                          constraints=[BoundConstraint("temperature_band", 0, 280.0, 321.0, penalty=-100.0)])
    w = whole.native
    w.reset_device(); w.track_extrema(True); w.rollout_device(K, N.POLICY_UNIFORM); torch.cuda.synchronize()
    ref = w.stats_dict()
    for k in ("steps", "episodes", "terminated", "truncated", "critical_shutdowns", "violations", "successes",
              "episode_length_sum", "violations_per_constraint"):
        assert summed[k] == ref[k], (k, summed[k], ref[k])
    assert abs(summed["return_sum"] - ref["return_sum"]) <= 1e-9 * abs(ref["return_sum"]) + 1e-6
    w.rollout_device(512, N.POLICY_UNIFORM); torch.cuda.synchronize()
    assert ext == w.read_extrema() and ext[0] is not None and ext[0] < ext[1], (ext, w.read_extrema())
    ws = w.get_state_host()[0]
    assert np.array_equal(np.concatenate(gathered).view(np.uint32), ws.view(np.uint32))
    assert summed["steps"] == n_total * K and summed["violations_per_constraint"][3] > 0
    print("NCCL_SHARDED_OK", world, summed["steps"], summed["violations_per_constraint"], flush=True)
dist.barrier()
dist.destroy_process_group()
