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


def fit_parameter_averaging(net, datasets, averaging_frequency: int = 1):
    """SparkComputationGraph.fit(rdd) with a ParameterAveragingTrainingMaster in ONE process (J:325-333, J:426; Python/gan.ipynb:177-187):
    every DataSet of the RDD goes to its own worker; each worker starts from the broadcast (parameters, updater state, iteration count), fits
    its minibatches -- at most `averaging_frequency` of them between two averagings -- and the driver then averages parameters AND updater
    state over the workers.  One native net plays the workers in turn (snapshot / restore), exactly what local[4] Spark does with model copies.
    `datasets`: list of workers, each a (x, y) pair or a list of (x, y) minibatches.  Returns the workers' last scores."""
    import numpy as np
    workers = [[d] if isinstance(d, tuple) else list(d) for d in datasets]
    if len(workers) == 1 and len(workers[0]) == 1:
        return [net.fit(*workers[0][0])]
    scores = [None] * len(workers)
    longest = max(len(w) for w in workers)
    for start in range(0, longest, max(1, averaging_frequency)):
        p0, s0, it0 = net.params(), net.updater_state(), net.iteration()
        psum, ssum, used, steps = np.zeros_like(p0, np.float64), np.zeros_like(s0, np.float64), 0, 0
        for wi, w in enumerate(workers):
            mine = w[start:start + max(1, averaging_frequency)]
            if not mine:
                continue
            net.set_params(p0); net.set_updater_state(s0); net.set_iteration(it0)
            for x, y in mine:
                scores[wi] = net.fit(x, y)
            psum += net.params(); ssum += net.updater_state(); used += 1; steps = max(steps, len(mine))
        net.set_params((psum / used).astype(np.float32)); net.set_updater_state((ssum / used).astype(np.float32)); net.set_iteration(it0 + steps)
    return scores
