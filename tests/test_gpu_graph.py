"""CUDA-graph capture of the step path: with the device-resident tick a captured sequence of single-step / fused
rollout launches replays with fresh random draws and walks exactly the trajectory of the same calls made eagerly."""
import numpy as np
import pytest

from util import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("kind", ["reactor", "grid", "robot"])
def test_graph_replay_matches_eager_stepping(mode, kind):
    """mode 1: the device tick is advanced by every launch; mode 2: base + per-launch sequence offsets, the captured
    sequence ends with commit_ticks() (and the single-step kernel draws its noise before the previous launch has finished)."""
    import torch
    import neorl_industrial as ni
    from neorl_industrial import _native as N
    n = 20_000
    dev = torch.device("cuda", 0)
    ek = {"reactor": N.ENV_CHEMICAL_REACTOR, "grid": N.ENV_POWER_GRID, "robot": N.ENV_ROBOT_ASSEMBLY}[kind]
    envs = [ni.NativeEnv(ek, n, device=0, seed=31) for _ in range(2)]
    graph_env, eager_env = envs
    for e in envs:
        e.reset_device()
    acts = torch.rand((graph_env.A, graph_env.pitch), device=dev) * 2 - 1
    bufs = [(e.empty(), e.empty(dtype=torch.uint8), e.empty(dtype=torch.uint8)) for e in envs]

    def body(e, b):
        e.step_device(acts, reward=b[0], flags=b[1], viol_mask=b[2])
        e.step_device(acts, reward=b[0], flags=b[1], viol_mask=b[2])
        e.rollout_device(16, N.POLICY_UNIFORM)
        e.step_device(acts, reward=b[0], flags=b[1], viol_mask=b[2])

    graph_env.use_device_tick(mode)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):           # warm-up outside capture (lazy allocations, function attributes)
        body(graph_env, bufs[0])
    torch.cuda.current_stream().wait_stream(side)
    graph_env.commit_ticks()                # (mode 2: the warm-up's launches; a no-op in mode 1)
    body(eager_env, bufs[1])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body(graph_env, bufs[0])
        graph_env.commit_ticks()
    body(eager_env, bufs[1])                # the capture itself does not execute: replay once for it
    g.replay()
    for _ in range(3):
        g.replay()
        body(eager_env, bufs[1])
    torch.cuda.synchronize()
    assert graph_env.tick == eager_env.tick == 5 * 19
    sg, stg, vg, dg = graph_env.get_state_host()
    se, ste, ve, de = eager_env.get_state_host()
    assert_bits_equal(sg, se, "state after graph replays vs eager")
    assert np.array_equal(stg, ste) and np.array_equal(vg, ve)
    assert_bits_equal(bufs[0][0][:n].cpu().numpy(), bufs[1][0][:n].cpu().numpy(), "last reward")
    cg, _ = graph_env.read_stats()
    ce, _ = eager_env.read_stats()
    assert cg[:8].tolist() == ce[:8].tolist()
    graph_env.use_device_tick(False)
    assert graph_env.tick == 5 * 19
    for e in envs:
        e.close()
