"""Connection records: the JSON ``Socket.log_pair`` prints for every pair of correlated packets
(examples/network_simple/network_simple_overrides.py:102-120,152-186), produced by the engine into HBM and compared
with the oracle record for record."""
import json

import numpy as np
import pytest

BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]
WINDOW = 8


def engine_run(backend, hostemu, seed, replica, horizon, cap=60000, builder="build", n=2, window=WINDOW, **kw):
    from itsim_b200 import netsimple
    sim = getattr(netsimple, builder)(connection_records=cap, connection_window=window, **kw)
    sim._lib = hostemu if backend == "hostemu" else None
    eng = sim.run_replicas(n, seed=seed, first_replica_id=replica)
    eng.run(horizon)
    eng.sim = sim
    return eng


def oracle_records(model):
    """(t_end, node, local ip, local port, inbound, remote ip, remote port, sent, received, t_start) tuples."""
    return list(model.connections)


def assert_same(got, ref):
    assert len(got) == len(ref)
    for name, col in (("node", 1), ("local_ip", 2), ("local_port", 3), ("inbound", 4), ("remote_ip", 5),
                      ("remote_port", 6), ("sent_bytes", 7), ("received_bytes", 8)):
        assert [int(x) for x in got[name]] == [int(r[col]) for r in ref], name
    assert np.array_equal(got["t_end"], np.array([r[0] for r in ref], dtype=np.float64))
    assert np.array_equal(got["t_start"], np.array([r[9] for r in ref], dtype=np.float64))


@pytest.mark.parametrize("backend", BACKENDS)
def test_records_match_the_oracle(backend, hostemu):
    import netsimple as oracle_netsimple
    seed, replica, horizon = 5, 1, 1500.0
    eng = engine_run(backend, hostemu, seed, replica, horizon)
    got = eng.connections(0)
    model = oracle_netsimple.run_replica(seed, replica, horizon, trace=False, conn_window=WINDOW)
    ref = oracle_records(model)      # before close(): unwinding the parked bodies flushes their sockets too
    model.close()
    assert eng.connections_total == len(ref) and len(ref) > 100
    assert_same(got, ref)
    # all three kinds of record occur: a completed pair, an unanswered outbound packet, an unanswered inbound one
    assert {int(h) for h in got["has"]} == {1, 2, 3}
    # the telemetry vector counts the records (user slot 6), and the run's other counters are those of a run
    # without the log
    from itsim_b200 import _lib as L, netsimple
    assert int(eng.replica_telemetry(0)[L.TM_USER0 + 6]) == len(ref)
    plain = netsimple.build()
    plain._lib = hostemu if backend == "hostemu" else None
    eng0 = plain.run_replicas(1, seed=seed, first_replica_id=replica)
    eng0.run(horizon)
    a, b = eng.replica_telemetry(0).copy(), eng0.replica_telemetry(0)
    a[L.TM_USER0 + 6] = 0
    assert (a == b).all()


@pytest.mark.parametrize("backend", BACKENDS)
def test_a_window_never_reached_gives_the_reference_records(backend, hostemu):
    """The reference's correlation dictionaries are unbounded; the engine keeps the newest `connection_window` entries
    per socket and direction and logs the one it gives up.  With a window the run never fills, the records are those
    of the unbounded dictionaries; a small window reports the same packets, some of them earlier and unpaired."""
    import netsimple as oracle_netsimple
    seed, replica, horizon = 2, 0, 1200.0
    unbounded = oracle_netsimple.run_replica(seed, replica, horizon, trace=False, conn_window=None)
    ref = oracle_records(unbounded)
    unbounded.close()
    eng = engine_run(backend, hostemu, seed, replica, horizon, window=1024, n=1)
    assert_same(eng.connections(0), ref)
    small = engine_run(backend, hostemu, seed, replica, horizon, window=4, n=1)
    bounded = oracle_netsimple.run_replica(seed, replica, horizon, trace=False, conn_window=4)
    ref_small = oracle_records(bounded)
    bounded.close()
    assert len(ref_small) > len(ref)
    assert_same(small.connections(0), ref_small)


@pytest.mark.parametrize("backend", BACKENDS)
def test_malware_overlay_records_match_the_oracle(backend, hostemu):
    """The router's relay sends from the socket of its other attachment (malware_simple.py:101-107): every beacon
    supersedes the unanswered one before it on that socket."""
    import netsimple as oracle_netsimple
    seed, replica, horizon = 3, 4, 900.0
    eng = engine_run(backend, hostemu, seed, replica, horizon, builder="build_malware")
    model = oracle_netsimple.MalwareModel(seed, replica, trace=False, conn_window=WINDOW)
    model.run(horizon)
    ref = oracle_records(model)
    model.close()
    got = eng.connections(0).copy()
    # the engine numbers a machine's second attachment as a node of its own; the oracle logs the machine
    renumber = {eng.sim.router_wan.index: eng.sim.router.index, eng.sim.bear.index: model.bear.index}
    got["node"] = [renumber.get(int(n), int(n)) for n in got["node"]]
    assert_same(got, ref)
    bear = oracle_netsimple.MalwareModel.BAD_NEWS_BEAR_ADDR
    assert sum(1 for r in ref if r[5] == bear and r[4] == 0) > 10


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("builder", ["build", "build_malware"])
def test_retired_wakeups_are_correlated_in_the_delivery_pass(backend, builder, hostemu):
    """With the log on, the daemons keep their fused receive loops (OP_RECV_FILTER_LOG): a wake-up the declared filter
    would reject is retired in the delivery pass, which then does that recv()'s correlation itself and logs under the
    counter value the wake-up would have had.  Records (after the read-side ordering) and every counter equal those of
    the same model assembled from plain receive loops, where each of these packets wakes its daemon."""
    fused = engine_run(backend, hostemu, 8, 40, 1800.0, builder=builder, n=3)
    plain = engine_run(backend, hostemu, 8, 40, 1800.0, builder=builder, n=3, fused=False)
    for r in range(3):
        a, b = fused.connections(r), plain.connections(r)
        assert fused.connections_total == plain.connections_total == len(a) > 1000
        assert np.array_equal(a, b), r
        assert np.array_equal(fused.replica_telemetry(r), plain.replica_telemetry(r)), r


@pytest.mark.parametrize("backend", BACKENDS)
def test_overflow_is_counted_not_written(backend, hostemu):
    eng = engine_run(backend, hostemu, 0, 0, 1500.0, cap=16)
    got = eng.connections(0)
    assert len(got) == 16 and eng.connections_total > 16


def test_a_model_assembled_without_the_logged_instructions_is_rejected(hostemu):
    from itsim_b200 import netsimple
    from itsim_b200._lib import EngineError
    sim = netsimple.build()
    sim.caps.update(connection_records=100, connection_window=WINDOW)
    sim._lib = hostemu
    with pytest.raises(EngineError, match="logged socket instructions"):
        sim.run_replicas(1)
    sim = netsimple.build(connection_records=100, connection_window=5000)
    sim._lib = hostemu
    with pytest.raises(EngineError, match="connection_window"):
        sim.run_replicas(1)


def test_documents_read_like_the_reference_log_lines(hostemu):
    """telemetry.connection_documents gives the dictionary `log_pair` dumps (same keys, same value conventions)."""
    from itsim_b200 import telemetry
    eng = engine_run("hostemu", hostemu, 5, 1, 900.0)
    docs = telemetry.connection_documents(eng.connections(0))
    keys = {"Connection_Type", "Local_IP", "Local_Port", "Direction", "Remote_IP", "Remote_Port", "Sent_Bytes",
            "Received_Bytes", "PID", "Start_Time", "End_Time", "Parent_Process_Path", "Parent_Process_Start_Time"}
    assert docs and all(set(d) == keys for d in docs)
    pair = next(d for d in docs if d["Local_IP"] != 0 and d["Remote_Port"] != 0)
    assert pair["Connection_Type"] == "UDP" and pair["Direction"] in ("Inbound", "Outbound")
    assert pair["Local_IP"].startswith("192.168.4.") and isinstance(pair["Local_Port"], str)
    assert pair["End_Time"] >= pair["Start_Time"] == pair["Parent_Process_Start_Time"]
    lonely = next(d for d in docs if d["Local_IP"] == 0)
    assert lonely["Local_Port"] == 0 and lonely["Received_Bytes"] == 0 and lonely["Direction"] == "Outbound"
    json.dumps(docs)


# ---- random streams against the live oracle (ITSB_SWEEP=N widens it) ---------------------------------------------------------
import os  # noqa: E402

SWEEP = int(os.environ.get("ITSB_SWEEP", "2"))


@pytest.mark.parametrize("i", range(SWEEP))
def test_records_of_a_random_stream_match_the_oracle(hostemu, i):
    """Connection records of a (seed, replica) stream no fixture holds -- 64-bit values included -- with a random
    correlation window: record for record, start and end times bit for bit."""
    import netsimple as oracle_netsimple
    rng = np.random.default_rng(4242 + i)
    seed = int(rng.integers(0, 1 << 63)) if i % 2 else int(rng.integers(0, 1 << 16))
    replica = int(rng.integers(0, 1 << 63)) if i % 3 == 0 else int(rng.integers(0, 1 << 20))
    horizon = float(rng.choice([700.0, 1100.0, 1600.0]))
    window = int(rng.choice([2, 4, 8, 64]))
    eng = engine_run("hostemu", hostemu, seed, replica, horizon, window=window, n=1)
    model = oracle_netsimple.run_replica(seed, replica, horizon, trace=False, conn_window=window)
    ref = oracle_records(model)
    model.close()
    got = eng.connections(0)
    assert eng.connections_total == len(ref)
    assert_same(got, ref)
