"""
Engine-side restatements of the reference's integration-test models for the shipped-``itsim`` primitives
(``/root/reference/tests/integ/*.py``).  Each function builds the same topology with ``itsim_b200.machine`` and
assembles the test's process bodies, citing the lines it follows.  ``oracle/machine_scenarios.py`` runs the
reference's own bodies for the same scenarios; ``tests/`` compare the two event for event.
"""
from . import ir
from .ir import (DST_BCAST, DST_PKT_SRC, DST_REGS, EXC, EXC_ITSIM_TIMEOUT, G1, R0, R1, R2, R3, SELF_TASK)
from .machine import Endpoint, Link, as_address
from .model import Simulator
from .random import constant, uniform
from .units import MS, MbPS, S

TAG_PING, TAG_PONG, TAG_CHATTER = 1, 2, 3
MARK_FAIL = 99


def _caps(sim: Simulator, trace_records: int) -> None:
    sim.caps.update(max_events=64, max_procs=128, max_sockets=64, arena_nodes=2048, trace_records=trace_records,
                    max_threads=64, max_tasks=64, max_signals=256)


def _ip(text: str) -> int:
    return int(as_address(text))


def packet_transfer(trace_records: int = 4096) -> Simulator:
    """tests/integ/test_packet_transfer.py:16-55."""
    sim = Simulator()
    a = sim.asm
    c_1s, c_5s = sim.const(1 * S), sim.const(5 * S)

    a.thread_program("client", 1)                       # :19-38
    a.bind(0)                                           # with context.node.bind(Protocol.UDP) as socket:
    a.li(R2, _ip("10.11.12.20"))
    a.li(R3, 9887)
    a.li(R0, 4)
    a.sendl(DST_REGS, R0, TAG_PING)                     #   socket.send(("10.11.12.20", 9887), 4, ping)
    a.recvt(c_1s, handler="unexpected_timeout")         #   packet = socket.recv(1 * S)
    a.recvt(c_5s, handler="second_recv")                #   try: socket.recv(5 * S); pytest.fail()
    a.mark(MARK_FAIL)
    a.label("second_recv")
    a.jnei(EXC, EXC_ITSIM_TIMEOUT, "cleanup_raise")     #   except Timeout: assert now() > 5.0
    a.mark(1)                                           #   finally: ledger.add("client")
    a.sclose()
    a.end_thread_program()
    a.label("unexpected_timeout")                       #   except Timeout: pytest.fail(...)
    a.mark(MARK_FAIL)
    a.label("cleanup_raise")
    a.sclose()
    a.reraise()

    a.thread_program("server", 2)                       # :41-47
    a.bind(9887)
    a.recvt()                                           #   packet = socket.recv()
    a.li(R0, 8)
    a.sendl(DST_PKT_SRC, R0, TAG_PONG)                  #   socket.send(packet.source, 8, pong)
    a.mark(2)                                           #   ledger.add("server")
    a.sclose()
    a.end_thread_program()

    link = Link("10.11.12.0/24", latency=uniform(100 * MS, 200 * MS), bandwidth=constant(100 * MbPS))
    pinger = Endpoint().connected_to_static(link, "10.11.12.10")
    pinger.run_proc_in(sim, 0.1, "client")
    ponger = Endpoint().connected_to_static(link, "10.11.12.20")
    ponger.run_proc(sim, "server")
    _caps(sim, trace_records)
    return sim


def broadcast(trace_records: int = 65536) -> Simulator:
    """tests/integ/test_broadcast.py:17-53: 20 endpoints on a /16, each broadcasting 128 B every U(3, 6) s."""
    sim = Simulator()
    a = sim.asm
    d_interval = sim.dist(uniform(3.0 * S, 6.0 * S))

    a.thread_program("server", 1)                       # :28-32
    a.bind(10000)
    a.label("loop")
    a.recvt()
    a.jmp("loop")
    a.end_thread_program()

    a.thread_program("client", 2)                       # :34-40
    a.bind(0)
    a.label("loop")
    a.advance(d_interval)
    a.li(R0, 128)
    a.sendl(DST_BCAST, R0, TAG_CHATTER, port=10000)     # for every non-loopback interface (there is one)
    a.jmp("loop")
    a.end_thread_program()

    link = Link("10.11.0.0/16", uniform(100 * MS, 200 * MS), constant(100 * MbPS))
    for n in range(100, 120):
        ep = Endpoint()
        ep.run_proc(sim, "server")                      # EndpointChattering.__init__ starts both before connecting
        ep.run_proc(sim, "client")
        ep.connected_to_static(link, as_address(n, link.cidr))
    _caps(sim, trace_records)
    return sim


def thread_lifecycle(trace_records: int = 4096) -> Simulator:
    """tests/integ/test_thread_lifecycle.py:21-84; known answer: marks 1..6 in order, final clock 120."""
    sim = Simulator()
    a = sim.asm
    c0, c5, c20, c100, c1000 = (sim.const(v) for v in (0.0, 5.0, 20.0, 100.0, 1000.0))

    a.thread_program("parent", 1)                       # :56-76
    a.mov(R0, SELF_TASK)
    a.fork("watcher", c0)                               # context.process.fork_exec(watcher, cemetary, context.process)
    a.clone("elder", c0, rd=R1)
    a.clone("cadet", c0, rd=R2)
    a.clone("super_long", c0, rd=R3)
    a.join(R1)
    a.join(R2)
    a.join(R3, c100, handler="timed_out")               # try: join(100); pytest.fail()
    a.mark(MARK_FAIL)
    a.label("timed_out")
    a.jnei(EXC, EXC_ITSIM_TIMEOUT, "__exit_raise")      # except Timeout: pass
    a.mark(4)                                           # cemetary.append("parent")
    a.pkill(SELF_TASK)                                  # context.process.kill()
    a.end_thread_program()

    a.thread_program("watcher", 2)                      # :21-23
    a.pwait(R0)
    a.mark(6)
    a.end_thread_program()

    a.thread_program("elder", 3)                        # :32-36
    a.clone("grandchild", c0, rd=R1)
    a.join(R1)
    a.mark(3)
    a.end_thread_program()

    a.thread_program("cadet", 4)                        # :39-42
    a.advance_const(c5)
    a.mark(1)
    a.end_thread_program()

    a.thread_program("super_long", 5)                   # :45-53
    a.advance_const(c1000, handler="finally")
    a.mark(MARK_FAIL)
    a.label("finally")                                  # except ThreadKilled: raise / finally: append
    a.mark(5)
    a.reraise()
    a.end_thread_program()

    a.thread_program("grandchild", 6)                   # :26-29
    a.advance_const(c20)
    a.mark(2)
    a.end_thread_program()

    Endpoint().with_proc_in(sim, 0, "parent")
    _caps(sim, trace_records)
    return sim


def process_lifecycle(trace_records: int = 4096) -> Simulator:
    """tests/integ/test_process_lifecycle.py:41-112."""
    sim = Simulator()
    a = sim.asm
    c0, c130, c730, c912, c1000, c2000, c100000 = (sim.const(v) for v in (0.0, 130.0, 730.0, 912.0, 1000.0, 2000.0,
                                                                        100000.0))
    a.thread_program("adameve", 1)                      # :84-101
    a.fork("cain", c0, rd=R0)                           # child[c.name] = fork_exec_in(c.year_born, c.life, ...)
    a.fork("abel", c0, rd=R1)
    a.gset(1, R1)                                       # the shared dict: Cain looks Abel up later
    a.fork("seth", c130, rd=R2)
    a.fork("humanity", c0, rd=R3)
    a.pwait(R1)                                         # for c in [Abel(), Cain(), Seth()]: child[c.name].wait()
    a.pwait(R0)
    a.pwait(R2)
    a.pwait(R3, c2000, handler="timed_out")             # try: child["Humanity"].wait(2000); pytest.fail()
    a.mark(MARK_FAIL)
    a.label("timed_out")
    a.jnei(EXC, EXC_ITSIM_TIMEOUT, "__exit_raise")
    a.mark(5)                                           # cemetary.add("adameve")
    a.end_thread_program()

    def life(name: str, mark: int, body) -> None:       # Character.life :41-46: try: _live() finally: add(name)
        a.thread_program(name, 2)
        body()
        a.mark(mark)
        a.end_thread_program()
        a.label("finally")
        a.mark(mark)
        a.reraise()

    def cain() -> None:                                 # :54-57
        a.now_sub(c130)
        a.advance_x2(handler="finally")                 # advance(Abel.YEAR_DEAD - now())
        a.pkill(G1)                                     # child["Abel"].kill()
        a.now_sub(c730)
        a.advance_x2(handler="finally")                 # advance(self.year_dead - now())

    life("cain", 1, cain)
    life("abel", 2, lambda: a.advance_const(c1000, handler="finally"))      # :67-68
    life("seth", 3, lambda: a.advance_const(c912, handler="finally"))       # :76-77
    life("humanity", 4, lambda: a.advance_const(c100000, handler="finally"))  # :85-86

    Endpoint().with_proc(sim, "adameve")
    _caps(sim, trace_records)
    return sim


def link_delay(trace_records: int = 64) -> Simulator:
    """tests/network/test_link.py:17-61: 1000 B at latency 1 s, bandwidth 8000 arrive after 2.0 s."""
    sim = Simulator()
    a = sim.asm
    a.thread_program("sender", 1)
    a.bind(9887)
    a.li(R2, _ip("192.168.1.78"))
    a.li(R3, 25001)
    a.li(R0, 1000)
    a.sendl(DST_REGS, R0, 0)
    a.end_thread_program()
    a.thread_program("receiver", 2)
    a.bind(25001)
    a.recvt()
    a.mark(1)
    a.end_thread_program()
    link = Link("192.168.1.0/24", latency=constant(1 * S), bandwidth=constant(8 * 1000 / (1 * S)))
    Endpoint().connected_to_static(link, "192.168.1.4").run_proc(sim, "sender")
    Endpoint().connected_to_static(link, "192.168.1.78").run_proc(sim, "receiver")
    _caps(sim, trace_records)
    return sim


SCENARIOS = {"packet_transfer": packet_transfer, "broadcast": broadcast, "thread_lifecycle": thread_lifecycle,
             "process_lifecycle": process_lifecycle, "link_delay": link_delay}
