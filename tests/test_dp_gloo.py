"""CPU test of the N>1 path: two gloo ranks, flat gradient arena all-reduce + averaging."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monopsr_b200.core import dp
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    scale = dp.allreduce_flat(g)
    mean = g * scale
    expect = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = bool(torch.allclose(mean, expect)) and dp.shard_samples(8, rank, world) == list(range(rank, 8, world))
    # two buckets of one arena, the first one asynchronous (the engine's head / towers split)
    a = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    w = dp.allreduce_start(a[600:])
    dp.allreduce_flat(a[:600])
    dp.allreduce_finish(w)
    ok = ok and bool(torch.allclose(a / world, expect))
    out[rank] = ok
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_single_process_is_identity():
    from monopsr_b200.core import dp
    g = torch.ones(10)
    assert dp.allreduce_flat(g) == 1.0 and float(g.sum()) == 10.0
    assert dp.allreduce_start(g) is None
    dp.allreduce_finish(None)
