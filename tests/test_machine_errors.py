"""
Edge cases and error paths of the shipped-itsim primitives, as the reference's unit tests pin them
(/root/reference/tests/machine/test_node.py, tests/network/test_link.py): ephemeral port sequence and wrap-around,
port collisions, route resolution, unknown addresses, recv time-outs, sockets closed under a receiver, negative delays.
Each case is a small model assembled here; it runs on the host emulation and (marked gpu) on the B200.
"""
import numpy as np
import pytest

from itsim_b200 import _lib as L
from itsim_b200 import ir
from itsim_b200.ir import DST_REGS, EXC, EXC_ITSIM_TIMEOUT, NONE, R0, R1, R2, R3
from itsim_b200.machine import Endpoint, Link, Relay, as_address
from itsim_b200.model import Simulator
from itsim_b200.random import constant

INF = float("inf")
BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def ip(text):
    return int(as_address(text))


def new_sim():
    sim = Simulator()
    sim.caps.update(max_events=64, max_procs=64, max_sockets=32, arena_nodes=512, trace_records=40000, max_threads=32,
                    max_tasks=32, max_signals=128)
    return sim


def run(sim, backend, hostemu, duration=INF):
    sim._lib = hostemu if backend == "hostemu" else None
    eng = sim.run_replicas(1, seed=0)
    eng.run(duration)
    return eng


def marks(eng):
    tr = eng.trace(0)
    return [(float(r["t"]), int(r["aux"])) for r in tr if r["kind"] == L.TRACE_MARK]


def status_of(sim, backend, hostemu, duration=INF):
    sim._lib = hostemu if backend == "hostemu" else None
    eng = sim.run_replicas(1, seed=0)
    with pytest.raises(L.EngineError) as err:
        eng.run(duration)
    return err.value.code, str(err.value)


@pytest.mark.parametrize("backend", BACKENDS)
def test_ephemeral_ports_count_up_and_wrap(backend, hostemu):
    """test_node.py:75-79: 32768, 32769, ..., 60999, then 32768 again."""
    sim = new_sim()
    a = sim.asm
    a.thread_program("binder", 1)
    a.li(R1, 0)
    a.label("loop")
    a.bind(0)
    a.getport(R0)
    a.jltu(R1, R2, "quiet")                 # report the first three and everything from the 28231st on
    a.mark_value(R0)
    a.label("quiet")
    a.jnei(R1, 2, "skip")
    a.li(R2, 28230)
    a.label("skip")
    a.sclose()
    a.addi(R1, R1, 1)
    a.li(R3, 28234)
    a.jltu(R1, R3, "loop")
    a.end_thread_program()
    Endpoint().connected_to_static(Link("10.0.0.0/24", constant(0.1), constant(1e6)), "10.0.0.1").run_proc(sim, "binder")
    got = [v for _, v in marks(run(sim, backend, hostemu))]
    assert got == [32768, 32769, 32770, 60998, 60999, 32768, 32769]


@pytest.mark.parametrize("backend", BACKENDS)
def test_port_already_in_use(backend, hostemu):
    """node.py:193-194 / test_node.py PortAlreadyInUse: the second bind of port 80 on the same node fails."""
    sim = new_sim()
    a = sim.asm
    a.thread_program("binder", 1)
    a.bind(80)
    a.advance_const(sim.const(10.0))
    a.end_thread_program()
    node = Endpoint().connected_to_static(Link("10.0.0.0/24", constant(0.1), constant(1e6)), "10.0.0.1")
    node.run_proc(sim, "binder")
    node.run_proc_in(sim, 1.0, "binder")
    code, text = status_of(sim, backend, hostemu)
    assert code == 3 and "0x50" in text


def two_links_model(dest, relay_cidr=None):
    """test_node.py:28-58 endpoint_2links: 192.168.1.0/24 and 10.10.128.0/17 with a Relay through 10.10.128.1."""
    sim = new_sim()
    a = sim.asm
    a.thread_program("sender", 1)
    a.bind(9887)
    a.li(R2, ip(dest))
    a.li(R3, 443)
    a.li(R0, 1234)
    a.sendl(DST_REGS, R0, 0)
    a.end_thread_program()
    a.thread_program("listener", 2)
    a.bind(443)
    a.recvt()
    a.mark(7)
    a.end_thread_program()
    small = Link("192.168.1.0/24", constant(0.5), constant(1e9))
    large = Link("10.10.128.0/17", constant(0.25), constant(1e9))
    me = Endpoint().connected_to_static(small, "192.168.1.4")
    if relay_cidr is None:
        me.connected_to_static(large, "10.10.192.1", [Relay("10.10.128.1")])
    else:
        me._interfaces[-1]._routes = [Relay("192.168.1.2", relay_cidr)]
    me.run_proc(sim, "sender")
    on_small = Endpoint().connected_to_static(small, "192.168.1.67")
    on_small.run_proc(sim, "listener")
    on_large = Endpoint().connected_to_static(large, "10.10.192.245")
    on_large.run_proc(sim, "listener")
    gateway = Endpoint().connected_to_static(large, "10.10.128.1")
    gateway.run_proc(sim, "listener")
    return sim


@pytest.mark.parametrize("backend", BACKENDS)
def test_solve_transfer_longest_prefix(backend, hostemu):
    """test_node.py:261-285: local on the small link, local on the large link, everything else through the relay (the
    gateway receives a packet that is not for it: transit, dropped -- node.py:245-250)."""
    for dest, node, latency, accepted in (("192.168.1.67", 1, 0.5, True), ("10.10.192.245", 2, 0.25, True),
                                          ("172.92.0.2", 3, 0.25, False)):
        eng = run(two_links_model(dest), backend, hostemu, 5.0)
        tr = eng.trace(0)
        deliveries = tr[tr["kind"] == L.EV_DELIVER]
        assert len(deliveries) == 1
        assert int(deliveries["node"][0]) == node
        assert bool(deliveries["aux"][0] & 0x80000000) == accepted
        assert deliveries["t"][0] == pytest.approx(latency, rel=1e-4)
        assert ([v for _, v in marks(eng)] == [7]) == accepted


@pytest.mark.parametrize("backend", BACKENDS)
def test_no_route_to_host_and_no_such_address(backend, hostemu):
    """test_node.py:287-290 NoRouteToHost; test_link.py:64-66 NoSuchAddress."""
    code, _ = status_of(two_links_model("172.99.0.2", relay_cidr="10.0.0.0/8"), backend, hostemu, 5.0)
    assert code == 2
    code, text = status_of(two_links_model("192.168.1.99"), backend, hostemu, 5.0)
    assert code == 1 and hex(ip("192.168.1.99")) in text


def recv_timeout_model(timeout, close_at=None):
    """test_node.py:229-258: a receiver with a time-out, a packet that arrives at t = 100."""
    sim = new_sim()
    a = sim.asm
    c_timeout = sim.const(timeout) if timeout is not None else ir.NO_TIMEOUT
    a.thread_program("receiver", 1)
    a.bind(80)
    if close_at is not None:
        a.clone("closer", sim.const(0.0))
    a.recvt(c_timeout, handler="timed_out")
    a.mark(1)                                   # SimulationResult.COMPLETE
    a.label("done")
    a.end_thread_program()
    a.label("timed_out")
    a.jnei(EXC, EXC_ITSIM_TIMEOUT, "__exit_raise")
    a.mark(2)                                   # except Timeout: SimulationResult.TIMEOUT
    a.jmp("done")
    a.thread_program("closer", 3)
    a.advance_const(sim.const(close_at or 0.0))
    a.sclose()
    a.end_thread_program()
    a.thread_program("sender", 2)
    a.bind(9887)
    a.advance_const(sim.const(99.0))
    a.li(R2, ip("10.0.0.1"))
    a.li(R3, 80)
    a.li(R0, 12345)
    a.sendl(DST_REGS, R0, 0)
    a.end_thread_program()
    link = Link("10.0.0.0/24", constant(1.0), constant(1e12))
    Endpoint().connected_to_static(link, "10.0.0.1").run_proc(sim, "receiver")
    Endpoint().connected_to_static(link, "10.0.0.2").run_proc(sim, "sender")
    return sim


@pytest.mark.parametrize("backend", BACKENDS)
def test_recv_timeouts(backend, hostemu):
    got = marks(run(recv_timeout_model(1000.0), backend, hostemu))
    assert [v for _, v in got] == [1] and got[0][0] == pytest.approx(100.0, rel=1e-6)
    got = marks(run(recv_timeout_model(50.0), backend, hostemu, 200.0))
    assert got == [(50.0, 2)]
    got = marks(run(recv_timeout_model(None), backend, hostemu))
    assert [v for _, v in got] == [1]


@pytest.mark.parametrize("backend", BACKENDS)
def test_socket_closed_under_a_receiver(backend, hostemu):
    """test_node.py:210-219: close() at t = 50 ends a blocked recv() with ValueError("Socket is closed"), which nobody
    catches: Simulator.run raises (here: replica status ITSB_EXC_UNCAUGHT carrying the exception code)."""
    code, text = status_of(recv_timeout_model(None, close_at=50.0), backend, hostemu)
    assert code == 7
    assert "0x3" in text.split("detail")[-1]


@pytest.mark.parametrize("backend", BACKENDS)
def test_negative_delay_is_a_value_error(backend, hostemu):
    sim = new_sim()
    a = sim.asm
    a.thread_program("bad", 1)
    a.advance_const(sim.const(-1.0))
    a.end_thread_program()
    Endpoint().connected_to_static(Link("10.0.0.0/24", constant(0.1), constant(1e6)), "10.0.0.1").run_proc(sim, "bad")
    code, _ = status_of(sim, backend, hostemu)
    assert code == 6
