"""
Model front-end: process bodies written as plain Python functions, compiled to engine programs
(itsim_b200/frontend.py) behind the reference's own module paths (itsim_b200/compat.py).

* bodies in tests/frontend_bodies.py, compiled, must give the golden traces that the reference's integration-test bodies
  produced on the oracle (tests/golden/machine_*.npz) -- event for event, on the host emulation and on the GPU;
* the host-side statements of the bodies (ledger / list updates, asserts on the clock) are replayed from the trace;
* where /root/reference exists, the reference's tests/integ/*.py run UNMODIFIED on the engine;
* constructs outside the subset are refused with a CompileError naming the line.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import load_golden_trace

HERE = os.path.dirname(os.path.abspath(__file__))
INF = float("inf")
REFERENCE_INTEG = "/root/reference/tests/integ"


def load_bodies(path=os.path.join(HERE, "frontend_bodies.py"), name="frontend_bodies_under_compat"):
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def non_marks(trace):
    from itsim_b200 import _lib as L
    return trace[trace["kind"] != L.TRACE_MARK]


def assert_same_events(got, ref, rtol=0.0):
    got, ref = non_marks(got), non_marks(ref)
    assert len(got) == len(ref), (len(got), len(ref))
    for field in ("kind", "prog", "node", "aux"):
        neq = np.nonzero(got[field] != ref[field])[0]
        assert len(neq) == 0, f"first {field} mismatch at event {neq[0]}: got {got[neq[0]]}, ref {ref[neq[0]]}"
    assert np.array_equal(got["t"], ref["t"]) if rtol == 0.0 else np.allclose(got["t"], ref["t"], rtol=rtol, atol=0.0)


def named_marks(sim, trace, names):
    """(time, golden code) of the MARK records whose host statement is one of ``names`` (text -> golden code)."""
    from itsim_b200 import _lib as L
    out = []
    for rec in trace[trace["kind"] == L.TRACE_MARK]:
        text = sim.frontend.effect_text(int(rec["aux"]))
        if text in names:
            out.append((float(rec["t"]), names[text]))
    return out


def golden_marks(ref):
    from itsim_b200 import _lib as L
    return [(float(r["t"]), int(r["aux"])) for r in ref[ref["kind"] == L.TRACE_MARK]]


def scenario(name, lib, seed=0, replica=0):
    """Builds one scenario from tests/frontend_bodies.py under the compatibility modules; returns (sim, module, marks)."""
    from itsim_b200 import compat
    with compat.installed(lib):
        B = load_bodies()
        sim = compat.Simulator()
        sim.seed, sim.replica = seed, replica
        if name == "packet_transfer":
            sim.frontend.program_ids.update(pinger=1, ponger=2)
            B.nodes = B.build_ping_pong(sim)
            marks = {"seen.add('pinger')": 1, "seen.add('ponger')": 2}
        elif name == "broadcast":
            sim.frontend.program_ids.update(listen=1, shout=2)
            B.nodes = B.build_chatter(sim)
            marks = {}
        elif name == "thread_lifecycle":
            sim.frontend.program_ids.update(root=1, observer=2, first=3, second=4, sleeper=5, leaf=6)
            B.log = []
            B.build_threads(sim, B.log)
            marks = {f"log.append('{n}')": c for n, c in (("second", 1), ("leaf", 2), ("first", 3), ("root", 4),
                                                          ("sleeper", 5), ("observer", 6))}
        elif name == "process_lifecycle":
            sim.frontend.program_ids.update(ancestors=1, life=2)
            B.graves = set()
            B.build_family(sim, B.graves)
            marks = {"graves.add(self.name)": None, "graves.add('ancestors')": 5}
        elif name == "large_lan":
            sim.frontend.program_ids.update(lan_server=1, lan_client=2)
            B.build_lan(sim, 4, 16, compat.per_process, lambda: compat.Router(forwarding=True))
            marks = {}            # 128 processes: the capacities are sized by compat.Simulator itself
        elif name == "echo":
            sim.frontend.program_ids.update(answerer=1, asker=2)
            B.build_echo(sim)
            marks = {}
        elif name == "daemon":
            sim.frontend.program_ids.update(forward_recv=1, caller=2)
            B.build_daemon(sim)
            marks = {}
        else:
            raise KeyError(name)
        return sim, B, marks


CASES = [
    ("thread_lifecycle", "machine_thread_lifecycle.npz", 0, 0, INF),
    ("process_lifecycle", "machine_process_lifecycle.npz", 0, 0, 4000.0),
    ("packet_transfer", "machine_packet_transfer_s0_r0.npz", 0, 0, 10.0),
    ("packet_transfer", "machine_packet_transfer_s5_r77.npz", 5, 77, 10.0),
    ("broadcast", "machine_broadcast_s0_r0.npz", 0, 0, 10.0),
    ("broadcast", "machine_broadcast_s3_r11_d60.npz", 3, 11, 60.0),
    ("large_lan", "machine_large_lan_s1_r2_d20_s4_p16.npz", 1, 2, 20.0),       # one program for all 64 clients
]


def check_case(name, golden, seed, replica, duration, lib):
    sim, B, marks = scenario(name, lib, seed, replica)
    sim.run(duration)
    got, ref = sim._engine.trace(0), load_golden_trace(golden)
    assert_same_events(got, ref)
    if name == "process_lifecycle":
        # the four `life` bodies share one statement text; the set they fill tells them apart
        assert B.graves == {"killer", "victim", "third", "ancestors"}
        assert [t for t, _ in golden_marks(ref)] == [t for t, _ in named_marks(sim, got, marks)]
    elif marks:
        assert named_marks(sim, got, marks) == golden_marks(ref)
    if name == "large_lan":
        assert [n for n, _ in sim.asm.entries] == ["lan_server", "lan_client", "__balk", "__wait_one"]   # shared programs
    if name == "packet_transfer":
        assert B.seen == {"pinger", "ponger"}
        assert len(B.clock_at_timeout) == 1 and B.clock_at_timeout[0] > 5.0      # replayed with now() = the record's time
    if name == "thread_lifecycle":
        assert B.log == ["second", "leaf", "first", "root", "sleeper", "observer"]
        assert sim.now() == 120.0
    if name == "broadcast" and duration >= 10.0:
        everyone = {n.address for node in B.nodes for n in node.interfaces() if not n.address.is_loopback}
        assert len(everyone) == 20
        for node in B.nodes:
            assert node.heard_from == everyone          # filled on the host from (MARK, value) record pairs
    sim._engine.close()


@pytest.mark.parametrize("name,golden,seed,replica,duration", CASES)
def test_compiled_bodies_give_the_reference_traces(hostemu, name, golden, seed, replica, duration):
    check_case(name, golden, seed, replica, duration, hostemu)


@pytest.mark.gpu
@pytest.mark.parametrize("name,golden,seed,replica,duration", CASES)
def test_compiled_bodies_give_the_reference_traces_on_the_gpu(name, golden, seed, replica, duration):
    check_case(name, golden, seed, replica, duration, None)


# The same source file executed on the reference's own itsim (oracle/machine_scenarios.frontend_scenario, goldens made by
# `python oracle/make_golden.py --frontend`): trace and the host-side state the bodies leave behind.
SAME_FILE = [("packet_transfer", "ping_pong_s0_r0"), ("packet_transfer", "ping_pong_s9_r4"), ("broadcast", "chatter_s1_r5"),
             ("thread_lifecycle", "threads_s0_r0"), ("process_lifecycle", "family_s0_r0"), ("large_lan", "lan_s2_r3"),
             ("echo", "echo_s0_r0"), ("echo", "echo_s6_r21"), ("daemon", "daemon_s0_r0"), ("daemon", "daemon_s11_r2")]


def check_same_file(name, tag, lib):
    import json
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "frontend_summaries.json")) as fh:
        want = json.load(fh)[tag]
    seed, replica = int(tag.split("_s")[1].split("_")[0]), int(tag.split("_r")[1])
    sim, B, _ = scenario(name, lib, seed, replica)
    sim.run(INF if want["duration"] is None else want["duration"])
    assert_same_events(sim._engine.trace(0), load_golden_trace(f"frontend_{tag}.npz"))
    assert sim.now() == pytest.approx(want["now"], rel=1e-9)
    state = want["state"]
    if name == "packet_transfer":
        assert sorted(B.seen) == state["seen"]
        assert B.clock_at_timeout == pytest.approx(state["clock_at_timeout"], rel=1e-9)
    elif name == "broadcast":
        assert [sorted(int(a) for a in n.heard_from) for n in B.nodes] == state["heard_from"]
    elif name == "thread_lifecycle":
        assert B.log == state["log"]
    elif name == "process_lifecycle":
        assert sorted(B.graves) == state["graves"]
    elif name == "echo":
        assert B.echoed == state["echoed"] and B.unanswered == state["unanswered"]
    elif name == "daemon":
        assert B.served == state["served"] and len(B.served) == 6
    sim._engine.close()


@pytest.mark.parametrize("name,tag", SAME_FILE)
def test_same_source_file_on_the_reference_and_on_the_engine(hostemu, name, tag):
    check_same_file(name, tag, hostemu)


@pytest.mark.gpu
@pytest.mark.parametrize("name,tag", SAME_FILE)
def test_same_source_file_on_the_reference_and_on_the_gpu(name, tag):
    check_same_file(name, tag, None)


@pytest.mark.skipif(not os.path.isdir("/root/reference/itsim"), reason="reference tree not on this box")
def test_goldens_are_what_the_reference_produces_now():
    import json
    import machine_scenarios as ms
    from conftest import GOLDEN, assert_trace_equal
    arr, out = ms.frontend_scenario("echo", 6, 21, 10.0)
    assert_trace_equal(arr, load_golden_trace("frontend_echo_s6_r21.npz"), rtol=0.0)
    with open(os.path.join(GOLDEN, "frontend_summaries.json")) as fh:
        assert json.load(fh)["echo_s6_r21"]["state"] == out["state"]


@pytest.mark.parametrize("backend", [pytest.param("hostemu", id="hostemu"),
                                     pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)])
def test_config_c4_compiled_from_python_equals_the_assembled_model(backend, hostemu):
    """The 1 024-endpoint LAN (BASELINE config 4) with its two bodies compiled from tests/frontend_bodies.py -- one
    program for all 1 024 clients, the peer address a per-process argument -- against machine_models.large_lan, whose
    bodies are written with the assembler: every counter of every replica."""
    from itsim_b200 import compat, machine_models
    lib = hostemu if backend == "hostemu" else None
    n, horizon = (2, 5.0) if backend == "hostemu" else (12, 6.0)
    with compat.installed(lib):
        B = load_bodies()
        sim = compat.Simulator()
        sim.frontend.program_ids.update(lan_server=1, lan_client=2)
        B.build_lan(sim, 16, 64, compat.per_process, lambda: compat.Router(forwarding=True))
        sim.caps["trace_records"] = 0
        compiled = sim.run_replicas(n, seed=4, first_replica_id=30)
        events = compiled.run(horizon)
    assert [name for name, _ in sim.asm.entries] == ["lan_server", "lan_client", "__balk", "__wait_one"]
    ref = machine_models.large_lan()
    ref._lib = lib
    assembled = ref.run_replicas(n, seed=4, first_replica_id=30)
    assert assembled.run(horizon) == events > 20000 * n
    for r in range(n):
        assert np.array_equal(compiled.replica_telemetry(r), assembled.replica_telemetry(r)), r
    assert compiled.state_bytes()["hot"] > 232448          # the in-place kernel's territory
    compiled.close()
    assembled.close()


@pytest.mark.gpu
def test_compiled_bodies_batched_on_the_gpu_match_the_host_emulation(hostemu):
    """The batched entry on a compiled model: 64 replicas of the chatter scenario, every counter of every replica."""
    engines = []
    for lib in (None, hostemu):
        sim, _, _ = scenario("broadcast", lib, seed=2, replica=0)
        sim.caps["trace_records"] = 0
        eng = sim.run_replicas(64, seed=2, first_replica_id=500)
        eng.run(90.0)
        engines.append(eng)
    for r in range(64):
        assert np.array_equal(engines[0].replica_telemetry(r), engines[1].replica_telemetry(r)), r
    for eng in engines:
        eng.close()


@pytest.mark.skipif(not os.path.isdir(REFERENCE_INTEG), reason="reference tree not on this box")
@pytest.mark.parametrize("module,test,golden,ids", [
    ("test_packet_transfer", "test_packet_transfer", "machine_packet_transfer_s0_r0.npz", dict(client=1, server=2)),
    ("test_broadcast", "test_broadcast", "machine_broadcast_s0_r0.npz", dict(server=1, client=2)),
    ("test_thread_lifecycle", "test_context_thread_lifecycle", "machine_thread_lifecycle.npz",
     dict(parent=1, watcher=2, elder=3, cadet=4, super_long=5, grandchild=6)),
    ("test_process_lifecycle", "test_context_process_lifecycle", "machine_process_lifecycle.npz",
     dict(adameve=1, life=2)),
])
def test_reference_integration_tests_run_unmodified_on_the_engine(hostemu, module, test, golden, ids):
    """The reference's own test function -- model construction, bodies, ``sim.run`` and its final asserts -- executed
    as is, with ``itsim`` / ``greensim`` resolving to the compatibility modules; the trace of the run equals the golden
    trace the same bodies produced on the reference's own classes (oracle/machine_scenarios.py)."""
    from itsim_b200 import compat
    made = []

    class Recording(compat.Simulator):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.frontend.program_ids.update(ids)
            made.append(self)

    with compat.installed(hostemu):
        sys.modules["itsim.simulator"].Simulator = Recording
        T = load_bodies(os.path.join(REFERENCE_INTEG, module + ".py"), "reference_" + module)
        getattr(T, test)()                       # raises if any of the reference's asserts fails
    assert len(made) == 1
    assert_same_events(made[0]._engine.trace(0), load_golden_trace(golden))
    made[0]._engine.close()


# ---- what the front-end refuses ----------------------------------------------------------------------------------------
def _compile(body, *args):
    from itsim_b200 import compat
    sim = compat.Simulator()
    compat.Endpoint().run_proc(sim, body, *args)
    return sim.compile()


def test_unsupported_constructs_are_refused_with_the_line(hostemu):
    from itsim_b200 import compat
    from itsim_b200.frontend import CompileError
    with compat.installed(hostemu):
        B = load_bodies(os.path.join(HERE, "frontend_refused.py"), "frontend_refused_under_compat")
        for body, needle in B.REFUSED:
            with pytest.raises(CompileError) as info:
                _compile(body)
            assert needle in str(info.value), (body.__name__, str(info.value))
            line = int(str(info.value).split("line ")[1].split(":")[0])
            first, count = B.first_line(body), len(__import__("inspect").getsourcelines(body)[0])
            assert first <= line < first + count, (body.__name__, str(info.value))


def test_interface_idiom_needs_single_homed_nodes(hostemu):
    from itsim_b200 import compat
    from itsim_b200.frontend import CompileError
    with compat.installed(hostemu):
        B = load_bodies()
        sim = compat.Simulator()
        one = B.Link("10.1.0.0/24", B.constant(0.1), B.constant(1000.0))
        two = B.Link("10.2.0.0/24", B.constant(0.1), B.constant(1000.0))
        node = compat.Router().connected_to_static(one, 1).connected_to_static(two, 1)
        node.run_proc(sim, B.lan_client, compat.per_process("10.2.0.9"))
        with pytest.raises(CompileError, match="one non-loopback interface"):
            sim.compile()


def test_compat_modules_are_removed_again(hostemu):
    from itsim_b200 import compat
    before = {k for k in sys.modules if k.split(".")[0] in ("itsim", "greensim")}
    with compat.installed(hostemu):
        import itsim.simulator
        assert itsim.simulator.Simulator is compat.Simulator
        with pytest.raises(TypeError):
            itsim.simulator.now()
    assert {k for k in sys.modules if k.split(".")[0] in ("itsim", "greensim")} == before
