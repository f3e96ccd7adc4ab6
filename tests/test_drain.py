"""Streaming the record logs home (itsb_drain_begin / itsb_drain_end): every replica's `network_event` and connection
records, packed on the device and copied in bulk, drain after drain, against the oracle's records and against the
per-replica reads of an engine that never drains."""
import numpy as np
import pytest

from test_connections import WINDOW, assert_same, oracle_records

BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def build(backend, hostemu, conn_cap, net_cap, builder="build"):
    from itsim_b200 import netsimple
    sim = getattr(netsimple, builder)(connection_records=conn_cap, connection_window=WINDOW)
    sim.caps["net_event_records"] = net_cap
    sim._lib = hostemu if backend == "hostemu" else None
    return sim


def per_replica(parts, key, first, r):
    """Replica r's records over a sequence of drains, concatenated."""
    out = [d[key][int(d[first][r]):int(d[first][r + 1])] for d in parts]
    return np.concatenate(out) if out else np.zeros(0)


@pytest.mark.parametrize("backend", BACKENDS)
def test_drained_records_are_the_oracles(backend, hostemu):
    """Small per-replica logs (they would overflow several times over the run) drained after every slice: nothing is
    lost, and what comes home for a replica, in order, is the oracle's record list."""
    import netsimple as oracle_netsimple
    seed, first_id, n = 5, 0, 3
    slices = [600.0, 300.0] + [150.0] * 6           # 1 800 s in all
    sim = build(backend, hostemu, conn_cap=8192, net_cap=1024)
    eng = sim.run_replicas(n, seed=seed, first_replica_id=first_id)
    parts = []
    for k, dt in enumerate(slices):
        eng.run(dt)
        eng.drain_begin()                           # overlapped form: the next run is queued before the copy is awaited
        if k + 1 < len(slices):
            eng.run_async(0.0)                      # (a zero-length run: one extra STOP event per replica, no records)
        parts.append(eng.drain_end(copy=True))
        if k + 1 < len(slices):
            eng.sync()
    assert all(p["net_dropped"] == 0 and p["conn_dropped"] == 0 for p in parts)
    assert sum(p["d2h_bytes"] for p in parts) >= sum(32 * len(p["net"]) + 40 * len(p["conn"]) for p in parts)
    for r in (1, 2):
        model = oracle_netsimple.run_replica(seed, first_id + r, sum(slices), trace=False, conn_window=WINDOW)
        ref_conn, ref_net = oracle_records(model), list(model.net_events)
        model.close()
        conn = per_replica(parts, "conn", "conn_first", r)
        net = per_replica(parts, "net", "net_first", r)
        assert len(ref_conn) > 2 * 8192             # the run's records would not fit the log that was drained
        assert_same(conn, ref_conn)
        assert [int(x) for x in net["node"]] == [e[1] for e in ref_net]
        assert [int(x) for x in net["port"]] == [e[2] for e in ref_net]
        assert [int(x) for x in net["type"]] == [e[3] for e in ref_net]
        assert np.array_equal(net["t"], np.array([e[0] for e in ref_net], dtype=np.float64))
    # the engine's own count of records produced (user slots 6 / 7) = what came home
    from itsim_b200 import _lib as L
    tm = eng.telemetry()
    assert int(tm[L.TM_USER0 + 6]) == sum(len(p["conn"]) for p in parts)
    assert int(tm[L.TM_USER0 + 7]) == sum(len(p["net"]) for p in parts)
    eng.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_drain_equals_the_per_replica_reads_and_reports_losses(backend, hostemu):
    sim = build(backend, hostemu, conn_cap=60000, net_cap=4096)
    a = sim.run_replicas(4, seed=9)
    a.run(1200.0)
    d = a.drain(copy=True)
    sim2 = build(backend, hostemu, conn_cap=60000, net_cap=4096)
    b = sim2.run_replicas(4, seed=9)
    b.run(1200.0)
    for r in range(4):
        assert np.array_equal(per_replica([d], "conn", "conn_first", r), b.connections(r))
        assert np.array_equal(per_replica([d], "net", "net_first", r), b.net_events(r))
    # after a drain the per-replica logs are empty, and an immediate second drain brings nothing
    assert len(a.connections(0)) == 0 and len(a.net_events(0)) == 0
    again = a.drain()
    assert len(again["conn"]) == 0 and len(again["net"]) == 0 and int(again["conn_first"][-1]) == 0
    # a log that is too small loses records and says how many
    small = build(backend, hostemu, conn_cap=64, net_cap=16).run_replicas(2, seed=9)
    small.run(1200.0)
    lost = small.drain()
    assert len(lost["conn"]) == 2 * 64 and len(lost["net"]) == 2 * 16
    assert lost["conn_dropped"] == sum(len(b.connections(r)) for r in range(2)) - 128
    assert lost["net_dropped"] == sum(len(b.net_events(r)) for r in range(2)) - 32
    # two pending drains at most
    small.drain_begin()
    small.drain_begin()
    from itsim_b200 import _lib as L
    with pytest.raises(L.EngineError):
        small.drain_begin()
    small.drain_end()
    small.drain_end()
    with pytest.raises(L.EngineError):
        small.drain_end()
    for e in (a, b, small):
        e.close()


@pytest.mark.gpu
def test_drain_at_scale_one_replica_against_the_oracle():
    """20 000 replicas with full logging, drained after every 120 s slice (1 800 s spin-up, then five slices): nothing
    dropped, and a replica sampled from the middle of the batch equals the oracle after all those drains."""
    import netsimple as oracle_netsimple
    n, seed, r = 20000, 0, 11111
    sim = build("gpu", None, conn_cap=16384, net_cap=2048)
    sim.caps["arena_nodes"] = 12288
    eng = sim.run_replicas(n, seed=seed)
    parts = []
    for dt in [600.0, 300.0, 300.0, 300.0, 300.0] + [120.0] * 5:
        eng.run(dt)
        parts.append(eng.drain(copy=False))
        assert parts[-1]["conn_dropped"] == 0 and parts[-1]["net_dropped"] == 0
        parts[-1] = {k: (v[int(parts[-1][f][r]):int(parts[-1][f][r + 1])].copy(), None)[0] if k in ("conn", "net") else v
                     for k, v, f in (("conn", parts[-1]["conn"], "conn_first"), ("net", parts[-1]["net"], "net_first"))}
    model = oracle_netsimple.run_replica(seed, r, 2400.0, trace=False, conn_window=WINDOW)
    ref_conn, ref_net = oracle_records(model), list(model.net_events)
    model.close()
    assert_same(np.concatenate([p["conn"] for p in parts]), ref_conn)
    net = np.concatenate([p["net"] for p in parts])
    assert [int(x) for x in net["port"]] == [e[2] for e in ref_net]
    assert np.array_equal(net["t"], np.array([e[0] for e in ref_net], dtype=np.float64))
    eng.close()
