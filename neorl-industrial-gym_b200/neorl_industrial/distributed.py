"""Multi-GPU layer: envs are independent, so the path shards by env index with ZERO inter-GPU traffic while
stepping; the only collective is one sum all-reduce of the counter block at the end of a rollout
(SURVEY section 8e). One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

Random streams are keyed by the GLOBAL env id (= shard offset + local index), so any sharding reproduces the
unsharded trajectories bit-for-bit.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def world_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of ``n_total`` global env ids: returns (offset, count) of ``rank``.
    The first ``n_total % world_size`` ranks own one extra env; blocks tile [0, n_total) exactly."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, extra = divmod(int(n_total), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def make_sharded(env_id: str, n_total: int, *, rank: Optional[int] = None, world_size: Optional[int] = None,
                 local_rank: Optional[int] = None, **kwargs):
    """``ni.make`` for this rank's shard of a ``n_total``-env job (device = cuda:LOCAL_RANK)."""
    from .utils import make
    r, w, lr = world_info()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    local_rank = lr if local_rank is None else local_rank
    offset, count = shard_bounds(n_total, world_size, rank)
    return make(env_id, num_envs=count, env_id_offset=offset, device=f"cuda:{local_rank}", batched=True, **kwargs)


def allreduce_stats(counters, sums, group=None):
    """Sum the int64 counters and fp64 sums of every rank (one all-reduce each; integer sums are exact and
    order independent). Accepts numpy arrays (moved through CPU tensors: gloo) or torch tensors (in place: NCCL
    for CUDA tensors). Returns arrays of the input kind. No-op when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counters, sums
    if isinstance(counters, np.ndarray):
        c = torch.from_numpy(np.ascontiguousarray(counters, np.int64).copy())
        s = torch.from_numpy(np.ascontiguousarray(sums, np.float64).copy())
        if dist.get_backend(group) == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            c, s = c.to(dev), s.to(dev)
        dist.all_reduce(c, group=group)
        dist.all_reduce(s, group=group)
        return c.cpu().numpy(), s.cpu().numpy()
    dist.all_reduce(counters, group=group)
    dist.all_reduce(sums, group=group)
    return counters, sums


_own_comm = {}


def nccl_comm(device_index: int, group=None) -> int:
    """An ``ncclComm_t`` (as an int) over the ranks of ``group`` for ``nig_allreduce_stats``: a communicator of this
    module's own (rank 0's NCCL id is handed round through the torch process group -- any backend --, then
    ``ncclCommInitRank``), cached per (group, device). Being separate from torch's communicators, its collectives can be
    queued on any stream without interfering with torch's."""
    import ctypes as C
    import torch.distributed as dist
    from . import _native as N
    key = (id(group), int(device_index))
    if key not in _own_comm:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = [None]
        if rank == 0:
            buf = (C.c_char * 128)()
            N.check(N.lib().nig_nccl_unique_id(C.cast(buf, C.c_void_p)))
            ident = [bytes(buf)]
        dist.broadcast_object_list(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        comm = C.c_void_p()
        N.check(N.lib().nig_nccl_comm_init(C.byref(comm), world, C.c_char_p(ident[0]), rank, int(device_index)))
        _own_comm[key] = comm.value
    return _own_comm[key]


def allreduce_device_stats(native, group=None, stream=None):
    """In-place all-reduce of a NativeEnv's device stats block over the ranks (no host staging): int64 counters and fp64
    sums (SUM) and the return-extremum keys (MAX). On NCCL process groups this is ONE grouped launch through the C ABI
    (``nig_allreduce_stats``) on the current torch stream; otherwise one ``torch.distributed`` all-reduce per part."""
    import torch
    import torch.distributed as dist
    from . import _native as N
    view = native.stats_tensor()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            st = native._stream(stream)
            N.check(N.lib().nig_allreduce_stats(native._h, nccl_comm(native.device, group), st))
        else:
            dist.all_reduce(view[:24], group=group)
            dist.all_reduce(view[24:].view(torch.float64), group=group)
            dist.all_reduce(native.extrema_tensor(), op=dist.ReduceOp.MAX, group=group)
    return view


def allreduce_extrema(native, group=None):
    """(return_min, return_max) over the finished episodes of EVERY rank: one MAX all-reduce of the two order-preserving
    int64 keys on the device (NCCL), then decode. (None, None) if no rank finished an episode."""
    import torch.distributed as dist
    keys = native.extrema_tensor()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    import torch
    torch.cuda.synchronize(native.torch_device())
    return native.decode_extrema(keys.cpu().numpy())


def allreduce_extrema_keys(keys, group=None):
    """Host-side flavour of :func:`allreduce_extrema`: ``keys`` = the two order-preserving int64 keys of this rank
    (``NativeEnv.extrema_tensor().cpu()`` or any int64[2]); one MAX all-reduce on the process group's CPU path, then
    decode. Returns (return_min, return_max) or (None, None)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from .vector import NativeEnv
    k = torch.as_tensor(np.ascontiguousarray(keys, np.int64)).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
    return NativeEnv.decode_extrema(k.numpy())
