"""CPU tests of the multi-GPU layer's host logic (world_size 2, gloo): shard partition, the final stats
all-reduce, and the sharding invariant (shards keyed by global env id == the unsharded run) using the oracle as
the stand-in stepper."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neorl_industrial.distributed import allreduce_extrema_keys, allreduce_stats, shard_bounds


def test_shard_bounds_tile_exactly():
    for n in (1, 7, 64, 65536, 1_048_576, 1_000_003):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_total, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    off, cnt = shard_bounds(n_total, world, rank)
    env = O.OracleEnv(O.GRID, cnt, env_id0=off, seed=11, exp_mode=1)
    env.reset()
    rs = O.rollout(env, T, O.POLICY_UNIFORM, want_reward_sum=True)
    counters = np.zeros(24, np.int64); counters[:16] = env.stats
    sums = np.zeros(8, np.float64); sums[2] = float(rs.astype(np.float64).sum())
    counters, sums = allreduce_stats(counters, sums)
    # return extrema across ranks: each rank's two order-preserving keys (restated here the way the kernel builds them),
    # one MAX all-reduce, decode
    def key(x):
        b = np.array(x, np.float64).view(np.uint64)
        return ((np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))) >> np.uint64(1)).astype(np.int64)
    ret = rs.astype(np.float64)
    ext = allreduce_extrema_keys(np.array([key(-ret).max(), key(ret).max()], np.int64))
    q.put((rank, off, env.state.copy(), counters, sums, ext))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_unsharded():
    from oracle import oracle as O
    n_total, T, world = 1001, 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = O.OracleEnv(O.GRID, n_total, seed=11, exp_mode=1)
    whole.reset()
    rs = O.rollout(whole, T, O.POLICY_UNIFORM, want_reward_sum=True)
    states = np.concatenate([r[2] for r in results])
    assert np.array_equal(states.view(np.uint32), whole.state.view(np.uint32)), "sharded trajectories differ"
    for r in results:        # every rank holds the same global sums after the all-reduce
        assert r[3][:16].tolist() == whole.stats.tolist()
        np.testing.assert_allclose(r[4][2], float(rs.astype(np.float64).sum()), rtol=1e-12)
        np.testing.assert_allclose(r[5], [float(rs.min()), float(rs.max())], rtol=3e-16)
    assert whole.stats[1] > n_total      # grid episodes are short: many auto-resets happened


def test_allreduce_is_noop_without_process_group():
    c, s = np.arange(24, dtype=np.int64), np.ones(8)
    c2, s2 = allreduce_stats(c, s)
    assert c2 is c and s2 is s
