"""
Parity tests proper: the CUDA engine, through the C ABI, against the oracle's golden vectors and against
size-independent properties at batch sizes the oracle cannot reach.  Run with ``pytest -m gpu`` on a B200.
"""
import numpy as np
import pytest

from conftest import assert_trace_equal, load_golden_trace, telemetry_matches_summary

pytestmark = pytest.mark.gpu

# Every engine kernel against the same golden vectors: "warp" = a replica per warp, state staged in shared memory
# (k_run); "pool" = a replica per lane, state in place in HBM, replicas regrouped by the kind of their next event through
# rings in shared memory (k_run_pool: what itsb_run picks from 8 192 replicas up, i.e. the kernel bench.py times);
# "vote" = the same with the regrouping confined to a warp (k_run_vote).  ITSB_KERNEL is read when the replicas are
# created.
KERNELS = [pytest.param(("warp", "k_run"), id="k_run"), pytest.param(("pool", "k_run_pool"), id="k_run_pool"),
           pytest.param(("vote", "k_run_vote"), id="k_run_vote")]


@pytest.fixture(params=KERNELS)
def kernel(request, monkeypatch):
    which, name = request.param
    monkeypatch.setenv("ITSB_KERNEL", which)
    return name


def make(trace_records=0, **caps):
    from itsim_b200 import netsimple
    sim = netsimple.build(trace_records=trace_records)
    sim.caps.update(caps)
    return sim


def test_trace_matches_golden_inside_a_batch(kernel):
    """Replica 0 of seed 0, traced while 5 replicas run side by side in one launch."""
    ref = load_golden_trace("netsimple_trace_s0_r0_1800s.npz")
    eng = make(trace_records=len(ref) + 16).run_replicas(5, seed=0)
    assert eng.kernel_name() == kernel
    eng.run(1800.0)
    assert_trace_equal(eng.trace(0), ref)
    assert np.all(eng.now() == 1800.0)
    eng.close()


def test_other_stream_matches_golden(kernel):
    ref = load_golden_trace("netsimple_trace_s7_r3_900s.npz")
    eng = make(trace_records=len(ref) + 16).run_replicas(2, seed=7, first_replica_id=3)
    assert eng.kernel_name() == kernel
    eng.run(900.0)
    assert_trace_equal(eng.trace(0), ref)
    eng.close()


def test_live_oracle_trace_short(kernel):
    """The oracle itself, run here, on a stream no fixture covers."""
    import netsimple as oracle_netsimple
    model = oracle_netsimple.run_replica(42, 17, 400.0)
    ref = model.tracer.as_array()
    model.close()
    eng = make(trace_records=len(ref) + 16).run_replicas(1, seed=42, first_replica_id=17)
    assert eng.kernel_name() == kernel
    eng.run(400.0)
    assert_trace_equal(eng.trace(0), ref)
    eng.close()


def test_64_bit_seed_and_replica_ids_against_the_live_oracle(kernel):
    """Seed and replica ids with their high words set (they enter the Philox counter, oracle/philox.py:146-148): three
    neighbouring replicas of one batch, each event for event against the oracle run here."""
    import netsimple as oracle_netsimple
    seed, first, horizon = 0x9E3779B97F4A7C15, (1 << 40) + 12345, 500.0
    refs = []
    for r in range(3):
        model = oracle_netsimple.run_replica(seed, first + r, horizon)
        refs.append(model.tracer.as_array())
        model.close()
    eng = make(trace_records=max(len(t) for t in refs) + 16).run_replicas(3, seed=seed, first_replica_id=first)
    assert eng.kernel_name() == kernel
    eng.run(horizon)
    for r in range(3):
        assert_trace_equal(eng.trace(r), refs[r])
    eng.close()


def test_summaries_match_golden_in_one_batch(golden_summaries, kernel):
    """Seed 0, replicas 0..99999 would be config C2; here the batch is the 64 first ids plus id 99999, and every
    replica that has a golden summary must match it counter for counter after 3600 s."""
    eng = make().run_replicas(64, seed=0)
    assert eng.kernel_name() == kernel
    events = eng.run(3600.0)
    for r in (0, 1, 2, 3, 63):
        s = golden_summaries[f"0:{r}:3600"]
        telemetry_matches_summary(eng.replica_telemetry(r), s)
    tm = eng.telemetry()
    assert int(tm[:9].sum()) == events
    assert np.array_equal(tm, sum(eng.replica_telemetry(r) for r in range(64)))
    eng.close()
    for key in ("0:99999:3600", "1:0:3600", "123456789012:5:3600"):
        seed, replica, horizon = (int(x) for x in key.split(":"))
        eng = make().run_replicas(1, seed=seed, first_replica_id=replica)
        assert eng.run(float(horizon)) == golden_summaries[key]["events"]
        telemetry_matches_summary(eng.replica_telemetry(0), golden_summaries[key])
        eng.close()


def test_full_horizon_c1(golden_summaries, kernel):
    """Config C1: the reference example's own horizon, 12 h = 43 200 s, 3.8 M events, one replica."""
    s = golden_summaries["0:0:43200"]
    eng = make().run_replicas(1, seed=0)
    assert eng.kernel_name() == kernel
    assert eng.run(43200.0) == s["events"]
    telemetry_matches_summary(eng.replica_telemetry(0), s)
    eng.close()


def test_bench_sized_batch_on_the_default_kernel_matches_the_oracle(golden_summaries, monkeypatch):
    """What bench.py times, proven against the oracle: 50 000 replicas of seed 0 in one batch, NO kernel override, so
    itsb_run's own choice is k_run_pool.  (1) the bench's regime (no trace: wake-ups a declared filter rejects are
    retired in the delivery pass): after 3 600 s every replica that has a golden summary -- ids 0, 1, 2, 3, 63 and
    the last one, 49 999 -- equals it counter for counter; (2) the same batch traced: replicas 0 and 49 999, event for
    event and timestamp for timestamp, against the oracle's golden traces."""
    monkeypatch.delenv("ITSB_KERNEL", raising=False)
    n = 50000
    eng = make(arena_nodes=12288).run_replicas(n, seed=0)
    assert eng.kernel_name() == "k_run_pool"
    events = eng.run(3600.0)
    for r in (0, 1, 2, 3, 63, n - 1):
        telemetry_matches_summary(eng.replica_telemetry(r), golden_summaries[f"0:{r}:3600"])
    tm = eng.telemetry()
    assert int(tm[:9].sum()) == events
    assert np.all(eng.now() == 3600.0)
    eng.close()
    first, last = load_golden_trace("netsimple_trace_s0_r0_600s.npz"), load_golden_trace("netsimple_trace_s0_r49999_600s.npz")
    eng = make(trace_records=max(len(first), len(last)) + 16, arena_nodes=12288).run_replicas(n, seed=0)
    assert eng.kernel_name() == "k_run_pool"
    eng.run(600.0)
    assert_trace_equal(eng.trace(0), first)
    assert_trace_equal(eng.trace(n - 1), last)
    eng.close()


def test_rerun_continues(golden_summaries):
    s = golden_summaries["0:1:3600"]
    eng = make().run_replicas(1, seed=0, first_replica_id=1)
    total = sum(eng.run(1200.0) for _ in range(3))
    assert total == s["events"] + 2                      # two extra STOP events
    assert int(eng.replica_telemetry(0)[10]) == s["bytes_sent"]
    eng.close()


def test_queued_async_launches(golden_summaries):
    s = golden_summaries["0:2:3600"]
    eng = make().run_replicas(1, seed=0, first_replica_id=2)
    for _ in range(4):
        eng.run_async(900.0)
    assert eng.sync() == s["events"] + 3
    assert eng.last_run_ms() > 0.0
    eng.close()


def test_gpu_equals_host_emulation_on_a_batch(hostemu):
    """Warp width 32 against warp width 1 of the same source: every counter of every replica."""
    gpu = make().run_replicas(96, seed=11, first_replica_id=1000)
    gpu.run(1500.0)
    sim = make()
    sim._lib = hostemu
    emu = sim.run_replicas(96, seed=11, first_replica_id=1000)
    emu.run(1500.0)
    for r in range(96):
        assert np.array_equal(gpu.replica_telemetry(r), emu.replica_telemetry(r)), r
    gpu.close()


def test_partition_invariance_and_large_batch_properties():
    """20 000 replicas x 120 s: no replica fails, conservation laws hold, and sampled replicas equal the same ids
    run alone (a replica depends only on (seed, id))."""
    from itsim_b200 import _lib as L
    n = 20000
    big = make().run_replicas(n, seed=5)
    events = big.run(120.0)
    tm = big.telemetry()
    assert int(tm[:9].sum()) == events
    assert int(tm[L.TM_KIND0 + L.EV_STOP]) == n
    assert int(tm[L.TM_KIND0 + L.EV_PROC_START]) >= 148 * n
    assert int(tm[L.TM_KIND0 + L.EV_TX_BEGIN]) == int(tm[L.TM_PACKETS_SENT])
    assert int(tm[L.TM_KIND0 + L.EV_TX_END]) <= int(tm[L.TM_KIND0 + L.EV_TX_BEGIN])
    assert int(tm[L.TM_DELIVERED]) + int(tm[L.TM_DROPPED]) == int(tm[L.TM_KIND0 + L.EV_DELIVER])
    assert int(tm[L.TM_LAT0:L.TM_LAT0 + L.LAT_BINS].sum()) == int(tm[L.TM_KIND0 + L.EV_TX_BEGIN])
    for r in (0, 7777, n - 1):
        alone = make().run_replicas(1, seed=5, first_replica_id=r)
        alone.run(120.0)
        assert np.array_equal(alone.replica_telemetry(0), big.replica_telemetry(r)), r
        alone.close()
    big.close()


def test_capacity_overflow_is_reported():
    from itsim_b200 import _lib as L
    eng = make(arena_nodes=64).run_replicas(3, seed=0)
    with pytest.raises(L.EngineError) as err:
        eng.run(3600.0)
    assert err.value.code == -14
    eng.close()


def test_tiny_pool_spills_to_the_arena_and_stays_exact(golden_summaries):
    s = golden_summaries["0:3:3600"]
    eng = make(max_events=16).run_replicas(1, seed=0, first_replica_id=3)
    assert eng.run(3600.0) == s["events"]
    telemetry_matches_summary(eng.replica_telemetry(0), s)
    eng.close()


def test_fused_loop_head_and_retired_wakeups_change_nothing():
    from itsim_b200 import netsimple
    results = []
    for fused, trace in ((False, 0), (True, 400000), (True, 0)):
        eng = netsimple.build(trace_records=trace, fused=fused).run_replicas(3, seed=21, first_replica_id=40)
        events = eng.run(3000.0)
        results.append((events, [eng.replica_telemetry(r) for r in range(3)], eng.now()))
        eng.close()
    for other in results[1:]:
        assert other[0] == results[0][0]
        for r in range(3):
            assert np.array_equal(other[1][r], results[0][1][r])
        assert np.array_equal(other[2], results[0][2])


@pytest.mark.parametrize("variant", ["vote", "pool", "pool:0", "pool:1", "pool:3"])
def test_every_engine_kernel_gives_the_same_replicas(variant, monkeypatch):
    """The two designs (a replica per warp: k_run; a replica per lane: k_run_vote, the default from 50 000 replicas up)
    and the experimental variants advance a replica through the same events: every counter of every replica, and the
    clock, equal k_run's.  ITSB_KERNEL selects the kernel when the replicas are created."""
    n, horizon = 1500, 400.0
    engines = {}
    kernel = variant.split(":")[0]
    if ":" in variant:
        monkeypatch.setenv("ITSB_POOL_SHAPE", variant.split(":")[1])     # another instantiation of k_run_pool
    for which in ("warp", kernel):
        monkeypatch.setenv("ITSB_KERNEL", which)
        eng = make().run_replicas(n, seed=21, first_replica_id=1000)
        engines[which] = (eng, eng.run(horizon), eng.kernel_name())
    ref, events, name = engines["warp"]
    got, events_k, name_k = engines[kernel]
    assert name == "k_run" and name_k == "k_run_" + kernel
    assert events_k == events
    for r in range(0, n, 7):
        assert np.array_equal(got.replica_telemetry(r), ref.replica_telemetry(r)), r
    assert np.array_equal(got.now(), ref.now())
    assert np.array_equal(got.telemetry(), ref.telemetry())
    for eng, _, _ in engines.values():
        eng.close()


def test_the_default_kernel_depends_on_the_batch(monkeypatch):
    monkeypatch.delenv("ITSB_KERNEL", raising=False)
    small = make().run_replicas(100, seed=0)
    assert small.kernel_name() == "k_run"
    small.close()
    big = make(arena_nodes=2048).run_replicas(50000, seed=0)
    assert big.kernel_name() == "k_run_pool"
    big.close()
