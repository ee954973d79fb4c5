"""N > 1 path on CPU: world_size-2 gloo run of the sharding / gather plumbing used by bench.py and the rollout driver.
The per-instance "solver" is a deterministic stand-in (the CUDA solve needs a GPU); what is under test is that the shards
partition the batch and that the gathered results are those of a single-process run, in instance order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_solve(idx: torch.Tensor):
    obj = (idx.double() * 0.25 + 1.0) ** 2
    status = (idx % 7 == 3).int()
    iters = (10 + idx % 9).int()
    return obj, status, iters


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = pkg("sharding")
    lo, hi = sh.shard_bounds(total, rank, world)
    sizes = [b - a for a, b in (sh.shard_bounds(total, r, world) for r in range(world))]
    obj, status, iters = _fake_solve(torch.arange(lo, hi))
    allres = sh.gather_results(sh.pack_results(obj, status, iters), world, sizes)
    ag = sh.AsyncGather(world, sizes)                   # bench.py's overlapped gather: on CPU tensors the same exchange
    ag.submit(sh.pack_results(obj, status, iters))
    ag.submit(sh.pack_results(obj + 1.0, status, iters))
    got = ag.wait()
    assert len(got) == 2 and torch.equal(got[0], allres) and torch.equal(got[1][:, 0], allres[:, 0] + 1.0) and not ag.pending
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the max-over-ranks timing reduction of bench.py
    q.put((rank, lo, hi, allres.numpy(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 37])
def test_two_rank_sharding_and_gather(total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total      # the shards partition [0, total)
    obj, status, iters = _fake_solve(torch.arange(total))
    ref = torch.stack([obj, status.double(), iters.double()], dim=1).numpy()
    for r in res:
        assert np.array_equal(r[3], ref)                                         # every rank holds the full result, in order
        assert r[4] == 2.0


def test_shard_bounds_cover_every_world_size():
    sh = pkg("sharding")
    for total in (1, 7, 1024, 65536):
        for world in (1, 2, 4, 8):
            b = [sh.shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
