"""
Shipped-itsim primitives: the engine's restatement of the reference's integration-test models
(itsim_b200/machine_models.py) against golden traces produced by the reference's OWN test bodies running unmodified on
the reference's itsim package + the oracle's greensim shim (oracle/machine_scenarios.py).  The known answers the
reference asserts -- cemetary order and final clock 120 of test_thread_lifecycle.py:83-84, the 2.0 s arrival of
test_link.py:56-61 -- are checked directly as well.
"""
import os

import numpy as np
import pytest

from conftest import assert_trace_equal, load_golden_trace

INF = float("inf")
CASES = [
    ("thread_lifecycle", "machine_thread_lifecycle.npz", 0, 0, INF),
    ("process_lifecycle", "machine_process_lifecycle.npz", 0, 0, 4000.0),
    ("packet_transfer", "machine_packet_transfer_s0_r0.npz", 0, 0, 10.0),
    ("packet_transfer", "machine_packet_transfer_s5_r77.npz", 5, 77, 10.0),
    ("broadcast", "machine_broadcast_s0_r0.npz", 0, 0, 10.0),
    ("broadcast", "machine_broadcast_s3_r11_d60.npz", 3, 11, 60.0),
    ("dhcp_exchange", "machine_dhcp_exchange_s0_r0.npz", 0, 0, 10.0),       # reference DHCPClient / DHCPServer classes
    ("dhcp_exchange", "machine_dhcp_exchange_s4_r9_d100.npz", 4, 9, 100.0),
]


def run_scenario(name, seed, replica, duration, lib=None, n=1):
    from itsim_b200 import machine_models
    sim = machine_models.SCENARIOS[name]()
    sim._lib = lib
    eng = sim.run_replicas(n, seed=seed, first_replica_id=replica)
    eng.run(duration)
    return eng


@pytest.mark.parametrize("name,golden,seed,replica,duration", CASES)
def test_hostemu_matches_reference_bodies(hostemu, name, golden, seed, replica, duration):
    eng = run_scenario(name, seed, replica, duration, lib=hostemu)
    assert_trace_equal(eng.trace(0), load_golden_trace(golden))


@pytest.mark.parametrize("name,golden,seed,replica,duration", CASES)
def test_helper_hops_without_the_interpreter_change_nothing(hostemu_fast, name, golden, seed, replica, duration):
    """ITSB_HIDDEN_FAST (what k_run_in_place is compiled with): select()'s helper processes take their hops through
    `hidden_wait_one` instead of the interpreter; the traces stay the reference's."""
    eng = run_scenario(name, seed, replica, duration, lib=hostemu_fast)
    assert_trace_equal(eng.trace(0), load_golden_trace(golden))


def test_known_answers_of_the_reference(hostemu):
    from itsim_b200 import _lib as L
    eng = run_scenario("thread_lifecycle", 0, 0, INF, lib=hostemu)
    tr = eng.trace(0)
    marks = [int(r["aux"]) for r in tr if r["kind"] == L.TRACE_MARK]
    assert marks == [1, 2, 3, 4, 5, 6]              # ["cadet", "grandchild", "elder", "parent", "super_long", "watcher"]
    assert eng.now()[0] == 120.0                    # TIMEOUT_SUPERLONG + max(DELAY_GRANDCHILD, DELAY_CADET)
    eng = run_scenario("link_delay", 0, 0, 5.0, lib=hostemu)
    tr = eng.trace(0)
    arrival = [r["t"] for r in tr if r["kind"] == L.EV_DELIVER]
    assert arrival == [2.0]                         # DURATION_TRANSFER_TOTAL
    assert int(tr[tr["kind"] == L.EV_DELIVER]["aux"][0]) == (1000 | 0x80000000)
    eng = run_scenario("process_lifecycle", 0, 0, 4000.0, lib=hostemu)
    tr = eng.trace(0)
    marks = {int(r["aux"]): float(r["t"]) for r in tr if r["kind"] == L.TRACE_MARK}
    assert marks == {2: 130.0, 1: 730.0, 3: 1042.0, 5: 3042.0}     # Abel, Cain, Seth die; adameve gives up on Humanity
    # tests/integ/test_dhcp_exchange.py:38: the three clients end up with hosts 100, 101, 102 -- each ACK is unicast to
    # the address the client took, so an accepted unicast delivery per host proves it
    eng = run_scenario("dhcp_exchange", 0, 0, 10.0, lib=hostemu)
    tm = eng.replica_telemetry(0)
    tr = eng.trace(0)
    acks = [r for r in tr if r["kind"] == L.EV_DELIVER and r["node"] != 0 and r["aux"] & 0x80000000]
    unicast_acks = [r for i, r in enumerate(tr) if r["kind"] == L.EV_DELIVER and r["aux"] & 0x80000000 and r["node"] != 0
                    and not (i + 1 < len(tr) and tr[i + 1]["kind"] == L.EV_DELIVER and tr[i + 1]["t"] == r["t"])
                    and not (i > 0 and tr[i - 1]["kind"] == L.EV_DELIVER and tr[i - 1]["t"] == r["t"])]
    assert sorted(int(r["node"]) for r in unicast_acks) == [1, 2, 3]
    assert int(tm[L.TM_PACKETS_SENT]) == 12 and len(acks) >= 3      # 3 x (DISCOVER, OFFER, REQUEST, ACK)


@pytest.mark.skipif(not os.path.isdir("/root/reference/itsim"), reason="reference tree not on this box")
def test_golden_is_what_the_reference_bodies_produce_now():
    import machine_scenarios as ms
    arr, _ = ms.packet_transfer(seed=5, replica=77)
    assert_trace_equal(arr, load_golden_trace("machine_packet_transfer_s5_r77.npz"), rtol=0.0)
    arr, _ = ms.thread_lifecycle()
    assert_trace_equal(arr, load_golden_trace("machine_thread_lifecycle.npz"), rtol=0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name,golden,seed,replica,duration", CASES)
def test_gpu_matches_reference_bodies(name, golden, seed, replica, duration):
    eng = run_scenario(name, seed, replica, duration, n=3)
    assert_trace_equal(eng.trace(0), load_golden_trace(golden))
    eng.close()


@pytest.mark.gpu
def test_gpu_equals_hostemu_on_a_broadcast_batch(hostemu):
    gpu = run_scenario("broadcast", 9, 100, 120.0, n=40)
    emu = run_scenario("broadcast", 9, 100, 120.0, lib=hostemu, n=40)
    for r in range(40):
        assert np.array_equal(gpu.replica_telemetry(r), emu.replica_telemetry(r)), r
    assert np.all(gpu.now() == 120.0)
    gpu.close()


# ---- random streams against the reference's own classes run live -------------------------------------------------------------
SWEEP = int(os.environ.get("ITSB_SWEEP", "4"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/itsim"), reason="reference tree not on this box")
@pytest.mark.parametrize("i", range(SWEEP))
def test_random_stream_equals_the_reference_classes(hostemu, hostemu_fast, i):
    """The reference's Link / Node / Socket / DHCP classes (oracle/machine_scenarios.py runs them unmodified) on a
    (seed, replica) stream no fixture holds, 64-bit values included, against the host build of the device source --
    both ways the helper hops of select() are compiled (interpreter / ITSB_HIDDEN_FAST)."""
    import machine_scenarios as ms
    from itsim_b200 import machine_models
    rng = np.random.default_rng(777 + i)
    seed = int(rng.integers(0, 1 << 63)) if i % 2 else int(rng.integers(0, 1 << 16))
    replica = int(rng.integers(0, 1 << 63)) if i % 3 == 0 else int(rng.integers(0, 1 << 20))
    which = ("broadcast", "dhcp_exchange", "packet_transfer", "large_lan")[i % 4]
    if which == "broadcast":
        duration = float(rng.choice([20.0, 75.0, 150.0]))
        ref, _ = ms.broadcast(seed=seed, replica=replica, duration=duration)
    elif which == "dhcp_exchange":
        duration = float(rng.choice([10.0, 60.0, 200.0]))
        ref, _ = ms.dhcp_exchange(seed=seed, replica=replica, duration=duration)
    elif which == "packet_transfer":
        duration = 10.0
        ref, _ = ms.packet_transfer(seed=seed, replica=replica)
    else:
        duration = float(rng.choice([8.0, 14.0]))
        ref, _ = ms.large_lan(seed=seed, replica=replica, duration=duration, subnets=3, per_subnet=8)
    for lib in (hostemu, hostemu_fast):
        if which == "large_lan":
            sim = machine_models.large_lan(trace_records=len(ref) + 16, subnets=3, per_subnet=8)
            sim._lib = lib
            eng = sim.run_replicas(1, seed=seed, first_replica_id=replica)
            eng.run(duration)
        else:
            eng = run_scenario(which, seed, replica, duration, lib=lib)
        assert_trace_equal(eng.trace(0), ref)
