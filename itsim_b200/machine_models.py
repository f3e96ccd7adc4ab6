"""
Engine-side restatements of the reference's integration-test models for the shipped-``itsim`` primitives
(``/root/reference/tests/integ/*.py``).  Each function builds the same topology with ``itsim_b200.machine`` and
assembles the test's process bodies, citing the lines it follows.  ``oracle/machine_scenarios.py`` runs the
reference's own bodies for the same scenarios; ``tests/`` compare the two event for event.
"""
from . import ir
from .ir import (DST_BCAST, DST_PKT_SRC, DST_REGS, EXC, EXC_ITSIM_TIMEOUT, G1, G6, NODE_INDEX, PKT_DST_IP, PKT_F0,
                 PKT_F1, PKT_TAG, R0, R1, R2, R3, SELF_TASK, ZERO)
from .machine import Endpoint, Link, Router, as_address
from .random import normal, num_bytes
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


# DHCP message types (itsim/network/service/dhcp/__init__.py:17-22) and the server's tables in the model table
DHCP_DISCOVER, DHCP_OFFER, DHCP_REQUEST, DHCP_ACK = 1, 2, 3, 4
T_CURSOR, T_UNIQUE, T_ALLOC_ADDR, T_ALLOC_UNIQ, T_ALLOC_CONF, T_RESERVED = 0, 1, 8, 72, 136, 200
T_RESERVED_SLOTS = 256


def dhcp_exchange(trace_records: int = 8192, clients: int = 3) -> Simulator:
    """tests/integ/test_dhcp_exchange.py:17-38: DHCPServer(100, cidr, gateway) behind Router.networking_daemon(UDP, 67),
    `clients` endpoints at address 0 running DHCPClient.  Bodies: itsim/network/service/dhcp/server.py:97-206,
    client.py:58-174, itsim/machine/node.py:319-350 (forward_recv)."""
    sim = Simulator()
    a = sim.asm
    link = Link("10.1.128.0/18", uniform(100 * MS, 200 * MS), constant(100 * MbPS))
    num_host_first = 100
    upper_num_host = int(link.cidr.hostmask)                      # server.py:67
    num_hosts_max = upper_num_host - num_host_first               # server.py:73
    first_address = int(link.cidr.network_address) + num_host_first
    c0, c30 = sim.const(0.0), sim.const(30 * S)                   # RESERVATION_TIME
    d_size = sim.dist(num_bytes(normal(100.0, 30.0), header=240))  # client.py:47-49, server.py:73-75
    assert num_hosts_max < 65536 and clients < 60

    # ---- Node.run_networking_daemon (node.py:340-348) + DHCPServer.on_packet -----------------------------------
    a.thread_program("dhcp_listen_first", 1)
    a.bind(67, detached=True)                                     # new_sock = self.bind(protocol, port)
    a.jmp("/dhcp_listen.body")
    a.end_thread_program()

    a.thread_program("dhcp_listen", 1)
    a.label("body")                                               # forward_recv:
    a.recvt()                                                     #   packet = socket.recv()
    a.clone("dhcp_listen", c0)                                    #   context.thread.clone(forward_recv, socket)
    a.jeqi(PKT_TAG, DHCP_DISCOVER, "discover")                    #   daemon.trigger -> on_packet (server.py:97-132)
    a.jeqi(PKT_TAG, DHCP_REQUEST, "request")
    a.jmp("end")
    # _handle_discover (server.py:134-176)
    a.label("discover")
    a.mov(R0, PKT_F0)                                             # node_id
    a.jnei(PKT_F1, 0, "given")                                    # address_maybe (never sent by the reference client)
    a.li(R3, 0)
    a.label("scan")                                               # for _ in range(num_hosts_max):
    a.tload(R2, ZERO, T_CURSOR)                                   #   suggestion = next(self._seq_num_hosts)
    a.mov(R1, R2)
    a.addi(R2, R2, 1)
    a.jnei(R2, num_hosts_max, "nowrap")
    a.li(R2, 0)
    a.label("nowrap")
    a.tstore(R2, ZERO, T_CURSOR)
    a.jltu(R1, G6, "in_table")                                    #   (table holds T_RESERVED_SLOTS addresses)
    a.jmp("found")
    a.label("in_table")
    a.tload(R2, R1, T_RESERVED)                                   #   if suggestion not in self._reserved: break
    a.jeqi(R2, 0, "found")
    a.addi(R3, R3, 1)
    a.jnei(R3, num_hosts_max, "scan")
    a.jmp("end")                                                  # else: decline
    a.label("found")
    a.addi(R1, R1, first_address)
    a.label("have")
    a.tload(R2, ZERO, T_UNIQUE)                                   # _AddressAllocation(suggestion): fresh unique
    a.addi(R2, R2, 1)
    a.tstore(R2, ZERO, T_UNIQUE)
    a.tstore(R1, R0, T_ALLOC_ADDR)
    a.tstore(R2, R0, T_ALLOC_UNIQ)
    a.tstore(ZERO, R0, T_ALLOC_CONF)
    a.addi(R3, R1, -first_address)
    a.tstore(R2, R3, T_RESERVED)                                  # self._reserved[suggestion] = True
    a.drawi(R3, d_size)
    a.sendl(DST_BCAST, R3, DHCP_OFFER, port=68, f0=R0, f1=R1)     # OFFER to the broadcast address
    a.advance_const(c30)                                          # advance(self._reservation_time)
    a.tload(R3, R0, T_ALLOC_CONF)                                 # if not confirmed and unique unchanged: forget it
    a.jnei(R3, 0, "end")
    a.tload(R3, R0, T_ALLOC_UNIQ)
    a.jne(R3, R2, "end")
    a.tload(R1, R0, T_ALLOC_ADDR)
    a.addi(R3, R1, -first_address)
    a.tstore(ZERO, R3, T_RESERVED)
    a.tstore(ZERO, R0, T_ALLOC_ADDR)
    a.jmp("end")
    a.label("given")
    a.mov(R1, PKT_F1)
    a.jmp("have")
    # _handle_request (server.py:178-206)
    a.label("request")
    a.jeqi(PKT_F1, 0, "end")
    a.mov(R0, PKT_F0)
    a.mov(R1, PKT_F1)
    a.tload(R2, R0, T_ALLOC_ADDR)
    a.jeqi(R2, 0, "end")                                          # node_id not in self._address_allocation
    a.jne(R2, R1, "end")                                          # spurious REQUEST
    a.li(R3, 1)
    a.tstore(R3, R0, T_ALLOC_CONF)
    a.drawi(R3, d_size)
    a.gset(5, R3)
    a.mov(R2, R1)
    a.li(R3, 68)
    a.sendl(DST_REGS, ir.G0 + 5, DHCP_ACK, f0=R0, f1=R1)          # ACK, unicast to the address
    a.label("end")
    a.end_thread_program()

    # ---- DHCPClient.run_client (client.py:58-174) -----------------------------------------------------------------
    a.thread_program("dhcp_client", 2)
    a.li(R0, T_RESERVED_SLOTS)
    a.gset(6, R0)                                                 # (constant for the server's table bound)
    a.addi(R0, NODE_INDEX, 1)                                     # node_id = context.process.node.uuid
    for attempt in range(3):                                      # for _ in range(self._dhcp_client_retries):
        t = f"a{attempt}_"
        a.bind(68)                                                # with context.node.bind(UDP, client_port, pid):
        a.drawi(R1, d_size)                                       # _dhcp_discover
        a.sendl(DST_BCAST, R1, DHCP_DISCOVER, port=67, f0=R0)
        a.xset(c30)                                               # _dhcp_iter_responses(OFFER)
        a.label(t + "offers")
        a.jxle0(t + "failed")
        a.xnow()
        a.recvt_x(handler=t + "offer_exc")
        a.xelapse()
        a.jnei(PKT_TAG, DHCP_OFFER, t + "offers")
        a.jne(PKT_F0, R0, t + "offers")
        a.jeqi(PKT_F1, 0, t + "offers")                           # if Field.ADDRESS in packet.payload
        a.mov(R1, PKT_F1)                                         # _dhcp_request(address_proposed)
        a.drawi(R2, d_size)
        a.sendl(DST_BCAST, R2, DHCP_REQUEST, port=67, f0=R0, f1=R1)
        a.getaddr(R3)                                             # address_orig
        a.setaddr(R1)                                             # tentatively take the address
        a.xset(c30)                                               # _dhcp_iter_responses(ACK)
        a.label(t + "acks")
        a.jxle0(t + "unconfirmed")
        a.xnow()
        a.recvt_x(handler=t + "ack_exc")
        a.xelapse()
        a.jnei(PKT_TAG, DHCP_ACK, t + "acks")
        a.jne(PKT_F0, R0, t + "acks")
        a.jne(PKT_DST_IP, R1, t + "acks")                         # packet.dest == address_proposed
        a.sclose()
        a.jmp("done")                                             # return True -> break
        a.label(t + "ack_exc")
        a.jnei(EXC, EXC_ITSIM_TIMEOUT, "cleanup_raise")           # except Timeout: return (generator ends)
        a.label(t + "unconfirmed")
        a.setaddr(R3)                                             # self._interface.address = address_orig
        a.sclose()
        a.jmp(t + "next")
        a.label(t + "offer_exc")
        a.jnei(EXC, EXC_ITSIM_TIMEOUT, "cleanup_raise")
        a.label(t + "failed")
        a.sclose()
        a.label(t + "next")
    a.label("done")
    a.end_thread_program()
    a.label("cleanup_raise")
    a.sclose()
    a.reraise()

    router = Router().connected_to_static(link, "10.1.128.1")
    router.run_proc(sim, "dhcp_listen_first")
    for _ in range(clients):
        Endpoint().connected_to_dynamic(link, sim)                 # node.py:135-154 = test_dhcp_exchange.py:25-27
    _caps(sim, trace_records)
    sim.caps["table_words"] = T_RESERVED + T_RESERVED_SLOTS
    return sim


LAN_PORT, TAG_LAN_REQUEST, TAG_LAN_REPLY = 10000, 4, 5


def large_lan(trace_records: int = 0, subnets: int = 16, per_subnet: int = 64) -> Simulator:
    """Config C4: `subnets` /25 access links of `per_subnet` endpoints + one forwarding router each, routers joined by a
    backbone /24, Relay default routes; same topology and bodies as ``oracle/machine_scenarios.large_lan`` (which
    builds it from the reference's classes).  One replica is ~650 KB of state at 16 x 64: the engine advances it in
    place in HBM (``k_run_in_place``)."""
    from .machine import Relay
    from .random import normal
    from .units import GbPS
    sim = Simulator()
    a = sim.asm
    d_interval = sim.dist(uniform(3.0, 6.0))

    a.thread_program("server", 1)
    a.bind(LAN_PORT)
    a.label("loop")
    a.recvt()
    a.jnei(PKT_TAG, TAG_LAN_REQUEST, "loop")            # if packet.payload.get("content") == "request":
    a.li(R0, 32)
    a.sendl(DST_PKT_SRC, R0, TAG_LAN_REPLY)             #     socket.send(packet.source, 32, reply)
    a.jmp("loop")
    a.end_thread_program()

    a.thread_program("client", 2)                       # r2 = peer address, r3 = port (run_proc arguments)
    a.bind(0)
    a.label("loop")
    a.advance(d_interval)
    a.li(R0, 128)
    a.sendl(DST_BCAST, R0, TAG_CHATTER, port=LAN_PORT)  # every non-loopback interface (there is one)
    a.li(R0, 64)
    a.sendl(DST_REGS, R0, TAG_LAN_REQUEST)              # socket.send((peer, 10000), 64, request)
    a.jmp("loop")
    a.end_thread_program()

    backbone = Link("10.30.0.0/24", normal(5 * MS, 1 * MS), constant(1 * GbPS))
    access = [Link(f"10.20.{i}.0/25", normal(5 * MS, 1 * MS), constant(1 * GbPS)) for i in range(subnets)]
    routers = []
    for i in range(subnets):
        router = Router(forwarding=True)
        router.connected_to_static(access[i], 1)
        router.connected_to_static(backbone, i + 1, [Relay(as_address(j + 1, backbone.cidr), f"10.20.{j}.0/25")
                                                    for j in range(subnets) if j != i])
        routers.append(router)
    for i in range(subnets):
        for h in range(per_subnet):
            ep = Endpoint()
            ep.connected_to_static(access[i], h + 2, [Relay(as_address(1, access[i].cidr))])
            peer = int(as_address(h + 2, access[(i + 1) % subnets].cidr))
            ep.run_proc(sim, "server")
            ep.run_proc(sim, "client", 0, 0, peer, LAN_PORT)
    n_ep = subnets * per_subnet
    procs = 2 * n_ep + 2 * n_ep + 64                    # bodies + the two select helpers of every blocked recv
    sim.caps.update(max_events=max(64, n_ep + 64), max_procs=procs, max_sockets=2 * n_ep + 16,
                    arena_nodes=max(2048, 64 * n_ep), trace_records=trace_records, max_threads=2 * n_ep + 16,
                    max_tasks=2 * n_ep + 16, max_signals=min(0x3FFF, 9 * n_ep + 64))
    return sim


SCENARIOS = {"large_lan": large_lan, "dhcp_exchange": dhcp_exchange, "packet_transfer": packet_transfer, "broadcast": broadcast, "thread_lifecycle": thread_lifecycle,
             "process_lifecycle": process_lifecycle, "link_delay": link_delay}
