"""
The N > 1 path on CPU: two gloo ranks each advance their shard of the replica set (host emulation of the engine,
test infrastructure) and sum their telemetry vectors; the result must equal one rank running the whole set.  On GPUs
the same sum is ncclAllReduce inside itsb_allreduce_telemetry (exercised by bench.py --gpus N).
"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_covers_the_set_once():
    from itsim_b200.sharding import shard
    for n, world in ((10, 3), (100000, 8), (7, 8), (0, 2)):
        ranges = [shard(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0
        for (f0, c0), (f1, _) in zip(ranges, ranges[1:]):
            assert f0 + c0 == f1
        assert sum(c for _, c in ranges) == n
        assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
    with pytest.raises(ValueError):
        shard(4, 2, 2)


def _rank_main(rank, world, port, n_total, horizon, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from itsim_b200 import _lib, netsimple
    from itsim_b200.sharding import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = netsimple.build()
    sim._lib = _lib.bind(ctypes.CDLL(os.path.join(ROOT, "tests", "hostemu", "libitsb_hostemu.so")))
    first, count = shard(n_total, rank, world)
    eng = sim.run_replicas(count, seed=9, first_replica_id=first)
    events = eng.run(horizon)
    tm = torch.from_numpy(eng.telemetry().astype(np.int64))
    dist.all_reduce(tm, op=dist.ReduceOp.SUM)
    # the per-replica histogram of config 3 (itsb_histogram): each rank bins its shard, the bins are summed
    hist = torch.from_numpy(eng.histogram(_lib.tm_user(netsimple.STAT_QUERIES), 0, 2, 32).astype(np.int64))
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    ev = torch.tensor([events], dtype=torch.int64)
    dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((tm.numpy().astype(np.uint64), int(ev.item()), hist.numpy().astype(np.uint64)))
    dist.destroy_process_group()


def test_two_ranks_equal_one(built, hostemu):
    import torch.multiprocessing as mp
    from itsim_b200 import netsimple
    n_total, horizon = 7, 900.0
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, n_total, horizon, out)) for r in range(2)]
    for p in procs:
        p.start()
    total, events, hist = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sim = netsimple.build()
    sim._lib = hostemu
    eng = sim.run_replicas(n_total, seed=9)
    assert eng.run(horizon) == events
    assert np.array_equal(eng.telemetry(), total)
    from itsim_b200 import _lib as L
    whole = eng.histogram(L.tm_user(netsimple.STAT_QUERIES), 0, 2, 32)
    assert np.array_equal(whole, hist) and int(hist.sum()) == n_total
