"""
Config C4: the large LAN (16 access links x 64 endpoints + 16 forwarding routers on a backbone).  Golden data comes
from the reference's own Link / Endpoint / Router / Relay classes (oracle/machine_scenarios.large_lan); one replica is
~650 KB of state, so on the GPU it runs through the in-place kernel (k_run_in_place).
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_trace_equal, load_golden_trace

BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module")
def summaries():
    with open(os.path.join(GOLDEN, "machine_summaries.json")) as fh:
        return json.load(fh)


def run(backend, hostemu, subnets, per_subnet, seed, replica, duration, trace_records, n=1):
    from itsim_b200 import machine_models
    sim = machine_models.large_lan(trace_records=trace_records, subnets=subnets, per_subnet=per_subnet)
    sim._lib = hostemu if backend == "hostemu" else None
    eng = sim.run_replicas(n, seed=seed, first_replica_id=replica)
    events = eng.run(duration)
    return eng, events


@pytest.mark.parametrize("backend", BACKENDS)
def test_small_instance_trace(backend, hostemu):
    ref = load_golden_trace("machine_large_lan_s1_r2_d20_s4_p16.npz")
    eng, _ = run(backend, hostemu, 4, 16, 1, 2, 20.0, len(ref) + 16, n=2)
    assert_trace_equal(eng.trace(0), ref)


def test_small_instance_trace_with_fast_helper_hops(hostemu_fast):
    ref = load_golden_trace("machine_large_lan_s1_r2_d20_s4_p16.npz")
    eng, _ = run("hostemu", hostemu_fast, 4, 16, 1, 2, 20.0, len(ref) + 16, n=2)
    assert_trace_equal(eng.trace(0), ref)


@pytest.mark.parametrize("backend", BACKENDS)
def test_full_size_instance(backend, hostemu, summaries):
    """1 024 endpoints + 16 routers, 8 simulated seconds, 175 881 events: counts, final record time and the digest of
    the whole trace's integer columns equal the reference classes' run."""
    import evtrace
    from itsim_b200 import _lib as L
    s = summaries["large_lan_s0_r0_d8_s16_p64"]
    eng, events = run(backend, hostemu, 16, 64, 0, 0, 8.0, s["events"] + 16)
    assert events == s["events"]
    tm = eng.replica_telemetry(0)
    for k in range(1, L.NUM_KINDS):
        assert int(tm[L.TM_KIND0 + k]) == s["counts"][L.KIND_NAMES[k]]
    tr = eng.trace(0)
    assert evtrace.trace_digest(tr) == s["t_digest"]
    assert tr["t"][-1] == pytest.approx(s["t_last"], rel=1e-9)
    assert eng.state_bytes()["hot"] > 232448            # larger than one SM's shared memory: the in-place path


@pytest.mark.gpu
@pytest.mark.parametrize("which,name", [("warp", "k_run_in_place"), ("vote", "k_run_vote_hbm"), ("pool", "k_run_pool_hbm")])
def test_both_large_model_kernels_give_the_reference_traces(which, name, summaries, monkeypatch):
    """Config C4's two kernels -- a warp per replica (k_run_in_place) and a lane per replica (k_run_vote_hbm, what
    itsb_run picks from 4 096 replicas up) -- against the reference classes' run: the small instance event for event,
    the full-size one by counts and trace digest, traced inside a batch."""
    import evtrace
    from itsim_b200 import _lib as L
    monkeypatch.setenv("ITSB_KERNEL", which)
    ref = load_golden_trace("machine_large_lan_s1_r2_d20_s4_p16.npz")
    eng, _ = run("gpu", None, 4, 16, 1, 2, 20.0, len(ref) + 16, n=3)
    assert_trace_equal(eng.trace(0), ref)
    eng.close()
    s = summaries["large_lan_s0_r0_d8_s16_p64"]
    eng, _ = run("gpu", None, 16, 64, 0, 0, 8.0, s["events"] + 16, n=3)
    assert eng.kernel_name() == name
    tm = eng.replica_telemetry(0)
    for k in range(1, L.NUM_KINDS):
        assert int(tm[L.TM_KIND0 + k]) == s["counts"][L.KIND_NAMES[k]]
    assert evtrace.trace_digest(eng.trace(0)) == s["t_digest"]
    eng.close()


@pytest.mark.gpu
def test_large_batch_on_the_default_kernel_equals_hostemu(hostemu, monkeypatch):
    """4 200 replicas: itsb_run's own choice is the pooled lane-per-replica kernel; sampled replicas equal the host
    build."""
    monkeypatch.delenv("ITSB_KERNEL", raising=False)
    gpu, _ = run("gpu", None, 16, 64, 3, 0, 2.0, 0, n=4200)
    assert gpu.kernel_name() == "k_run_pool_hbm"
    for r in (0, 2100, 4199):
        emu, _ = run("hostemu", hostemu, 16, 64, 3, r, 2.0, 0, n=1)
        assert np.array_equal(gpu.replica_telemetry(r), emu.replica_telemetry(0)), r
    gpu.close()


@pytest.mark.gpu
def test_gpu_batch_equals_hostemu(hostemu):
    gpu, ev_gpu = run("gpu", hostemu, 16, 64, 7, 50, 6.0, 0, n=24)
    emu, ev_emu = run("hostemu", hostemu, 16, 64, 7, 50, 6.0, 0, n=24)
    assert ev_gpu == ev_emu
    for r in range(24):
        assert np.array_equal(gpu.replica_telemetry(r), emu.replica_telemetry(r)), r
    gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["vote", "pool"])
@pytest.mark.parametrize("subnets", [4, 8])
def test_mid_size_tables_on_the_lane_kernels(which, subnets, hostemu, monkeypatch):
    """Models whose staged tables need more than the 48 KB of shared memory a kernel gets without opting in (4 and 8
    subnets x 64 endpoints: 57 600 and 115 456 bytes) on the lane-per-replica kernels: the launch succeeds and every
    sampled replica equals the host build."""
    monkeypatch.setenv("ITSB_KERNEL", which)
    gpu, ev_gpu = run("gpu", None, subnets, 64, 5, 10, 2.0, 0, n=40)
    assert gpu.kernel_name().startswith("k_run_" + which), gpu.kernel_name()
    emu, ev_emu = run("hostemu", hostemu, subnets, 64, 5, 10, 2.0, 0, n=40)
    assert ev_gpu == ev_emu
    for r in (0, 13, 39):
        assert np.array_equal(gpu.replica_telemetry(r), emu.replica_telemetry(r)), r
    gpu.close()
