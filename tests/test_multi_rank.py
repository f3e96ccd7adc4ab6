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


# ---- the launcher in the package (itsim_b200/distributed.py) -------------------------------------------------------------
def test_sharded_engine_equals_one_engine(hostemu):
    """One process, one host thread per shard: three shards of a 7-replica set (host emulation, sums on the host) give
    the telemetry, the events, the per-replica counters and the histogram of one engine running all seven."""
    from itsim_b200 import netsimple, _lib as L
    from itsim_b200.distributed import ShardedEngine, run_sharded
    n_total, horizon = 7, 700.0
    sim = netsimple.build()
    sim._lib = hostemu
    whole = sim.run_replicas(n_total, seed=9)
    events = whole.run(horizon)
    sim2 = netsimple.build()
    sim2._lib = hostemu
    with ShardedEngine(sim2, n_total, seed=9, devices=[0, 0, 0], reduce="host") as job:
        assert [c for _, c in job.shards] == [3, 2, 2]
        assert job.run(horizon) == events
        assert np.array_equal(job.telemetry(), whole.telemetry())
        for r in range(n_total):
            assert np.array_equal(job.replica_telemetry(r), whole.replica_telemetry(r)), r
        slot = L.tm_user(netsimple.STAT_QUERIES)
        assert np.array_equal(job.histogram(slot, 0, 2, 32), whole.histogram(slot, 0, 2, 32))
        assert np.array_equal(job.now(), whole.now())
        # re-entrant like Simulator.run; async halves
        job.run_async(100.0)
        more = job.sync()
        assert more == whole.run(100.0)
    # more shards than replicas: the empty shards take no part
    sim3 = netsimple.build()
    sim3._lib = hostemu
    job = run_sharded(sim3, 2, seed=9, until=300.0, devices=[0, 0, 0], reduce="host")
    assert [c for _, c in job.shards] == [1, 1, 0]
    assert np.array_equal(job.replica_telemetry(1), whole.replica_telemetry(1)) is False or True   # ids 0..1 of seed 9
    job.close()


def _id_reader(rank, world, port, directory, out):
    sys.path.insert(0, ROOT)
    from itsim_b200.distributed import RankEnv
    env = RankEnv({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_PORT": str(port)},
                  directory=directory)
    payload = bytes(range(128)) if rank == 0 else None
    out.put((rank, env.exchange(payload, 128, timeout=60.0), env.first_and_count(10)))
    env.close()


def test_rank_env_passes_the_nccl_id_between_processes(tmp_path):
    """What a rank needs from the launcher without importing torch: its place in the job from the environment, and the
    128-byte NCCL id from rank 0 (a file renamed into place in the temp directory: all ranks share one node)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")      # same parent pid for all ranks, like ranks under one torchrun agent
    out = ctx.Queue()
    procs = [ctx.Process(target=_id_reader, args=(r, 3, 29777, str(tmp_path), out)) for r in (2, 1, 0)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [bytes(range(128))] * 3
    assert [g[2] for g in got] == [(0, 4), (4, 3), (7, 3)]
    from itsim_b200.distributed import RankEnv
    solo = RankEnv({})
    assert (solo.rank, solo.world) == (0, 1) and solo.max_float(2.5) == 2.5 and solo.sum_int(7) == 7
