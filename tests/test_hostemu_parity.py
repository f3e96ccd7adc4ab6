"""
Interpreter semantics without a GPU: the device source compiled at warp width 1 (tests/hostemu, test infrastructure)
against the oracle's golden vectors.  The GPU parity tests proper are in test_gpu_parity.py.
"""
import numpy as np
import pytest

from conftest import assert_trace_equal, load_golden_trace, telemetry_matches_summary


def make(hostemu, trace_records=0, **caps):
    from itsim_b200 import netsimple
    sim = netsimple.build(trace_records=trace_records)
    sim.caps.update(caps)
    sim._lib = hostemu
    return sim


def test_trace_matches_golden(hostemu):
    ref = load_golden_trace("netsimple_trace_s0_r0_1800s.npz")
    sim = make(hostemu, trace_records=len(ref) + 16)
    eng = sim.run_replicas(1, seed=0)
    eng.run(1800.0)
    assert_trace_equal(eng.trace(0), ref)
    assert eng.now()[0] == 1800.0


def test_other_stream_matches_golden(hostemu):
    ref = load_golden_trace("netsimple_trace_s7_r3_900s.npz")
    sim = make(hostemu, trace_records=len(ref) + 16)
    eng = sim.run_replicas(1, seed=7, first_replica_id=3)
    eng.run(900.0)
    assert_trace_equal(eng.trace(0), ref)


def test_summaries_match_golden(hostemu, golden_summaries):
    checked = 0
    for key, s in golden_summaries.items():
        seed, replica, horizon = (int(x) for x in key.split(":"))
        if horizon > 3600:
            continue
        sim = make(hostemu)
        eng = sim.run_replicas(1, seed=seed, first_replica_id=replica)
        events = eng.run(float(horizon))
        assert events == s["events"], key
        telemetry_matches_summary(eng.replica_telemetry(0), s)
        checked += 1
    assert checked >= 5


def test_rerun_continues(hostemu, golden_summaries):
    """Repeated run() continues like the reference (test_thread.py:148-152): 3 x 1200 s == 3600 s, apart from the
    extra STOP events and the seq numbers they take."""
    s = golden_summaries["0:1:3600"]
    sim = make(hostemu)
    eng = sim.run_replicas(1, seed=0, first_replica_id=1)
    total = sum(eng.run(1200.0) for _ in range(3))
    assert total == s["events"] + 2
    tm = eng.replica_telemetry(0)
    from itsim_b200 import _lib as L
    assert int(tm[L.TM_KIND0 + L.EV_STOP]) == 3
    assert int(tm[L.TM_BYTES_SENT]) == s["bytes_sent"]
    assert eng.now()[0] == 3600.0


def test_batch_is_partition_invariant(hostemu):
    """Replica r's trajectory depends only on (seed, r), not on which batch or position it runs in."""
    sim = make(hostemu)
    whole = sim.run_replicas(6, seed=3, first_replica_id=10)
    whole.run(600.0)
    part = make(hostemu).run_replicas(2, seed=3, first_replica_id=13)
    part.run(600.0)
    assert np.array_equal(whole.replica_telemetry(3), part.replica_telemetry(0))
    assert np.array_equal(whole.replica_telemetry(4), part.replica_telemetry(1))
    assert not np.array_equal(whole.replica_telemetry(0), whole.replica_telemetry(1))
    total = whole.telemetry()
    assert np.array_equal(total, sum(whole.replica_telemetry(i) for i in range(6)))


def test_capacity_overflow_is_reported_not_hidden(hostemu):
    from itsim_b200 import _lib as L
    sim = make(hostemu, arena_nodes=64)
    eng = sim.run_replicas(1, seed=0)
    with pytest.raises(L.EngineError) as err:
        eng.run(3600.0)
    assert err.value.code == -14


def test_tiny_pool_spills_to_the_arena_and_stays_exact(hostemu, golden_summaries):
    """The future-event pool is a cache: when it is full, events continue in the HBM arena and order is unchanged."""
    s = golden_summaries["0:3:3600"]
    sim = make(hostemu, max_events=16)
    eng = sim.run_replicas(1, seed=0, first_replica_id=3)
    assert eng.run(3600.0) == s["events"]
    telemetry_matches_summary(eng.replica_telemetry(0), s)


def test_fused_loop_head_and_retired_wakeups_change_nothing(hostemu):
    """Three encodings of the same model must agree counter for counter: plain instructions; the fused RECV_FILTER
    loop head with tracing on (every wake-up is an event of its own); and tracing off, where filtered wake-ups are
    retired inside the delivery pass."""
    from itsim_b200 import netsimple
    results = []
    for fused, trace in ((False, 0), (True, 400000), (True, 0)):
        sim = netsimple.build(trace_records=trace, fused=fused)
        sim._lib = hostemu
        eng = sim.run_replicas(3, seed=21, first_replica_id=40)
        events = eng.run(3000.0)
        results.append((events, [eng.replica_telemetry(r) for r in range(3)], eng.now()))
    for other in results[1:]:
        assert other[0] == results[0][0]
        for r in range(3):
            assert np.array_equal(other[1][r], results[0][1][r])
        assert np.array_equal(other[2], results[0][2])


# ---- random streams against the live oracle ---------------------------------------------------------------------------------
# ITSB_SWEEP=N widens the sweep (an offline run of N = 48 is recorded in DESIGN.md §2); the default keeps the suite short.
import os  # noqa: E402

SWEEP = int(os.environ.get("ITSB_SWEEP", "3"))


@pytest.mark.parametrize("i", range(SWEEP))
def test_random_stream_equals_the_live_oracle(hostemu, i):
    """(seed, replica id) drawn from all 64 bits of both -- the high words go into the Philox counter
    (oracle/philox.py:146-148) -- on one of the three override-path models; the whole trace, timestamps included."""
    import netsimple as oracle_netsimple
    from itsim_b200 import netsimple
    rng = np.random.default_rng(20261020 + i)
    seed = int(rng.integers(0, 1 << 63)) if i % 2 else int(rng.integers(0, 1 << 20))
    replica = int(rng.integers(0, 1 << 63)) if i % 3 else int(rng.integers(0, 100000))
    duration = float(rng.choice([400.0, 900.0, 1500.0]))
    which = ("plain", "malware", "backdoor")[i % 3]
    if which == "plain":
        model = oracle_netsimple.run_replica(seed, replica, duration)
        build = netsimple.build
    elif which == "malware":
        model = oracle_netsimple.MalwareModel(seed, replica)
        model.run(duration)
        build = netsimple.build_malware
    else:       # the overlay with the knobs of examples/backdoor/backdoor.py, gaps short enough to see beacons
        from greensim.random import expo as o_expo, uniform as o_uniform
        from itsim_b200.random import expo, num_bytes, uniform
        targets = netsimple.BACKDOOR_C2
        model = oracle_netsimple.MalwareModel(
            seed, replica, beacon_gap=o_uniform(2.1, 4.9), privileged_gap=o_uniform(30.0, 50.0), c2_targets=list(targets),
            beacon_size=oracle_netsimple.num_bytes(o_expo(1024.0), header=128, upper=20 * 1024 * 1024), infected_index=23)
        model.run(duration)

        def build(trace_records):
            return netsimple.build_malware(
                trace_records=trace_records, beacon_gap=uniform(2.1, 4.9), privileged_gap=uniform(30.0, 50.0),
                c2_targets=targets, beacon_size=num_bytes(expo(1024.0), header=128, upper=20 * 1024 * 1024),
                infected_index=23)
    ref = model.tracer.as_array()
    model.close()
    sim = build(trace_records=len(ref) + 16)
    sim._lib = hostemu
    eng = sim.run_replicas(1, seed=seed, first_replica_id=replica)
    eng.run(duration)
    assert_trace_equal(eng.trace(0), ref)


@pytest.mark.parametrize("i", range(max(1, SWEEP // 3)))
def test_untraced_random_stream_equals_the_live_oracle(hostemu, i):
    """Without a trace the delivery pass takes its short cuts (retired wake-ups, virtual queues of sleeping workstations,
    Bloom words): every counter of a random stream, run long enough for workstations to sleep and wake, against the
    oracle's own counts."""
    import netsimple as oracle_netsimple
    rng = np.random.default_rng(99 + i)
    seed = int(rng.integers(0, 1 << 63)) if i % 2 else int(rng.integers(0, 1 << 16))
    replica = int(rng.integers(0, 1 << 63)) if i % 3 == 0 else int(rng.integers(0, 1 << 20))
    horizon = float(rng.choice([1900.0, 2600.0, 3300.0]))
    model = oracle_netsimple.run_replica(seed, replica, horizon, trace=False)
    s = model.tracer.summary()
    s.update(model.stats)
    model.close()
    sim = make(hostemu)
    eng = sim.run_replicas(1, seed=seed, first_replica_id=replica)
    events = eng.run(horizon)
    assert events == s["events"]
    telemetry_matches_summary(eng.replica_telemetry(0), s)
