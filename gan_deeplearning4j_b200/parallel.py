"""Data-parallel plumbing: one process per GPU (torchrun), NCCL inside libb200gan.so for the gradient all-reduce.

Replaces SparkComputationGraph + ParameterAveragingTrainingMaster (J:325-333; Python/gan.ipynb:177-187): instead of
averaging parameters and updater state every <=10 local minibatches, every D / G update sums the gradient vector over
ranks (one ncclAllReduce over NVLink) and divides by the global minibatch inside the updater kernel.  The two coincide
for linear updaters at averagingFrequency=1 (tests/test_parallel_cpu.py); for Adam/RmsProp the all-reduce is what
north_star mandates.

torch.distributed is used only to carry the 128-byte NCCL unique id to the other ranks and for barriers.
"""
from __future__ import annotations

import os
from typing import Callable, Tuple


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_batch(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of the global minibatch for `rank` (weak scaling keeps stop-start fixed)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def exchange_unique_id(dist, rank: int, make_id: Callable[[], bytes]) -> bytes:
    """Rank 0 creates the NCCL unique id (b2g_comm_unique_id) and broadcasts it through torch.distributed."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad NCCL unique id")
    return bytes(uid)


def attach_communicator(ctx, dist, rank: int, world: int):
    """Give a b200gan Context its NCCL communicator; afterwards Net.fit / Gan.step all-reduce their gradients."""
    from .engine import comm_unique_id
    if world <= 1:
        return ctx
    ctx.comm_init(world, rank, exchange_unique_id(dist, rank, comm_unique_id))
    return ctx
