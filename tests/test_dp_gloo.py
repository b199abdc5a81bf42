"""CPU world_size-2 (gloo) test of the data-parallel host logic: shard ranges from the product's C ABI,
per-rank partial gradient/loss sums (computed here by the oracle, standing in for the GPU kernel),
one all-reduce of the flat [P+3] buffer, identical Adam on every rank — the protocol of DESIGN.md §4."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load
    from oracle import oracle as O
    P = load()
    ocfg = O.make_config(O.PDE_BURGERS, 8, 2, 1001, 77, 65)
    w = O.init_params(ocfg).astype(np.float64)
    xs = []
    for k in range(3):
        b, e = P.shard_range(O.set_count(ocfg, k), rank, world)     # the product's sharding function (host-only ABI call)
        xs.append(O.sample(ocfg, k, b, e))
    g, s = O.loss_grad(ocfg, w, *xs)                                  # this rank's shard contribution
    buf = torch.from_numpy(np.concatenate([g, s]))
    dist.all_reduce(buf)                                              # the single collective of a step
    m = np.zeros_like(w); v = np.zeros_like(w)
    O.adam(ocfg, w, m, v, buf.numpy()[:-3].copy(), 1)
    q.put((rank, buf.numpy().copy(), w.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_allreduce_equals_single_rank():
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ocfg = O.make_config(O.PDE_BURGERS, 8, 2, 1001, 77, 65)
    w = O.init_params(ocfg).astype(np.float64)
    g, s = O.loss_grad(ocfg, w, *[O.sample(ocfg, k) for k in range(3)])
    ref = np.concatenate([g, s])
    assert np.array_equal(res[0][1], res[1][1])                       # every rank sees the same reduced buffer
    assert np.abs(res[0][1] - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.array_equal(res[0][2], res[1][2])                       # replicated Adam keeps weights identical
