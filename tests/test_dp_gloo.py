"""Data-parallel host logic on CPU: two gloo ranks (127.0.0.1).  The flow's kernels need a GPU, so what
is checked here is everything around them: the flat gradient arena is averaged across ranks in place,
the per-parameter gradient views alias that arena, and the per-rank input sharding of bench.py differs
by rank.  (On the GPU box the same code runs over NCCL; the shard-vs-full gradient invariant itself is a
-m gpu test: tests/test_flow_gpu.py::test_full_size_properties.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nf_b200
from oracle import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)   # identical replicas, like DP after a broadcast
        model = nf_b200.RealNVP(9, 32, 16, 2, None)
        model.enable_data_parallel()
        assert model._dp_world == world
        # every rank fabricates a different "local gradient" in the arena layout
        dflat = torch.full((model._layout.total,), float(rank + 1))
        dflat[::7] = float(10 * (rank + 1))
        views = model._grad_views(dflat)
        model._allreduce_grads(dflat)
        expect = torch.full((model._layout.total,), sum(range(1, world + 1)) / world)
        expect[::7] = 10 * sum(range(1, world + 1)) / world
        ok = torch.allclose(dflat, expect)
        # views alias the arena, so parameters see the averaged gradient without a copy
        ok = ok and all(v.data_ptr() >= dflat.data_ptr() for v in views)
        p, off = model._trainable[5]
        ok = ok and torch.equal(views[5], dflat[off:off + p.numel()].view(p.shape))
        # replicas stay identical
        flat = model._flat.clone()
        dist.all_reduce(flat)
        ok = ok and torch.allclose(flat, model._flat * world)
        # the conditioning encoder averages its arena the same way
        enc = nf_b200.GEncoder(12, 6, hidden=16, n_hidden=2).enable_data_parallel()
        eflat = torch.full((enc._total,), float(rank + 1))
        eviews = enc._grad_views(eflat)
        enc._allreduce_grads(eflat)
        ok = ok and torch.allclose(eflat, torch.full((enc._total,), sum(range(1, world + 1)) / world))
        ok = ok and eviews[0].data_ptr() == eflat.data_ptr() + 4 * enc._slots[0][1]
        # rank-sharded synthetic inputs (bench.py uses seed 1234 + rank)
        x, _ = synth.make_inputs(8, 9, 16, seed=1234 + rank)
        xs = [torch.zeros(8, 9) for _ in range(world)]
        dist.all_gather(xs, torch.tensor(x))
        ok = ok and not torch.equal(xs[0], xs[1])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_arena_is_averaged_across_two_gloo_ranks():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
