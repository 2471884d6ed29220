"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, unique-id exchange, and the
semantics the CUDA path implements -- "all-reduce(sum) of per-rank gradient sums, divide by the global minibatch"."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dl4j_oracle as o


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gan_deeplearning4j_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.env_rank_world() == (rank, world, rank)
        uid = parallel.exchange_unique_id(dist, rank, lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # gradient all-reduce semantics on a BN-free net (BN uses per-replica batch statistics by design)
        mk = lambda: o.Net([o.Dense(6, 5, "tanh", updater=o.Adam(1e-2), l2=1e-3, name="a"), o.Dense(5, 4, "lrelu", 0.2, updater=o.Adam(1e-2), name="b"),
                            o.Output(4, 1, updater=o.Adam(1e-2), name="out")], seed=7)
        rng = np.random.default_rng(0)
        X = rng.standard_normal((8, 6)); Y = rng.uniform(0, 1, (8, 1))
        lo, hi = parallel.shard_batch(8, world, rank)
        net = mk()
        for step in range(3):
            net.compute_gradient_and_score(X[lo:hi], Y[lo:hi])
            g = torch.from_numpy(net.grads_flat().copy())
            dist.all_reduce(g, op=dist.ReduceOp.SUM)                         # the one collective per update
            grads, off = {}, 0
            for li, _, p, shape, order in net.param_table():
                k = int(np.prod(shape)); grads[(li, p)] = g.numpy()[off:off + k].reshape(shape, order=order.upper()); off += k
            net.apply_update(8, grads=grads)                                  # divide by the GLOBAL minibatch
        ref = mk()
        for step in range(3):
            ref.fit(X, Y)
        np.testing.assert_allclose(net.params_flat(), ref.params_flat(), rtol=1e-10, atol=1e-12)
        with pytest.raises(ValueError):
            parallel.shard_batch(7, world, rank)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_equals_big_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
