"""
The flagship workload: ``examples/network_simple`` (reference ``examples/network_simple/network_simple.py:42-387``
on the override classes of ``network_simple_overrides.py``), built on this package's modelling API and compiled for
the engine.  ``build()`` mirrors ``init()`` (network_simple.py:311-387); each ``_prog_*`` restates one process body
as an IR program, citing the lines it follows.  RNG draw order per replica follows SURVEY.md Appendix B.
"""
from ipaddress import ip_address, ip_network
from typing import Optional, Sequence, Tuple

from . import ir
from .ir import (DST_BCAST, DST_ENT, DST_PKT_SRC, DST_SELF, NODE_ADDR, NODE_HOST, PKT_F0, PKT_F1, PKT_SIZE, PKT_TAG,
                 R0, R1, R2, R3, SELF)
from .ir import G0, R0, R1, R2, R3  # noqa: F811
from .ir import PKT_F0 as _PKT_F0, PKT_F1 as _PKT_F1, DST_REGS
from .model import DrawnNode, Endpoint, Network, Simulator, Workstation
from .random import Dist, bounded, constant, distribution_index, expo, normal, num_bytes, uniform
from .tags import Tag
from .units import B, GbPS, H, MB, MIN, MS, MbPS, S, US

# program ids reported in trace records (shared with oracle/netsimple.py)
PROG_DHCP_SERVE, PROG_BLINKING, PROG_CLIENT, PROG_MDNS, PROG_LLMNR = 1, 2, 3, 4, 5
PROG_QUERY_MDNS, PROG_QUERY_LLMNR, PROG_TIMEOUT_HELPER, PROG_TX, PROG_DELIVER = 6, 7, 8, 9, 10
PROG_CALL_HOME, PROG_NO_GOOD, PROG_SUDO_TRANSMIT, PROG_EXFILTRATE, PROG_ROUTE, PROG_C2 = 11, 12, 13, 14, 15, 16

# payload "content" values (network_simple.py:135-149,152-167,190-221)
TAG_DISCOVER, TAG_OFFER, TAG_REQUEST, TAG_ACK, TAG_QUERY, TAG_RESOLVE, TAG_IAM = 1, 2, 3, 4, 5, 6, 7
TAG_EXFIL = 8

EXC_COMPLETE = 1          # Workstation._Complete (network_simple.py:78-79): a plain Exception, not an Interrupt
STAT_QUERIES, STAT_RESOLVED, STAT_TIMEOUTS = 0, 1, 2
STAT_BEACONS, STAT_ROUTED, STAT_C2_RECEIVED = 3, 4, 5
STAT_FIRST_BEACON_US, STAT_EXFIL_BYTES = 8, 9      # slots 6 and 7 count the connection / network_event records

BAD_NEWS_BEAR_ADDR = (1 << 24) | (2 << 16) | (3 << 8) | 4     # malware_simple.py:16-18
BAD_NEWS_BEAR_PORT = 666
LOCAL_ROUTER_PORT = 777

NUM_ADDRESSES_RESERVED = 2
MIN_NUM_ENDPOINTS = 3


def _programs(sim: Simulator, num_workstations: int, fused: bool) -> None:
    a = sim.asm
    d_awake = sim.dist(expo(2.0 * H))                                         # network_simple.py:101
    d_sleeping = sim.dist(expo(10.0 * MIN))                                   # :102
    d_ready = sim.dist(expo(5.0 * US))                                        # :103
    d_dhcp = sim.dist(num_bytes(normal(100.0 * B, 30.0 * B), header=240 * B))  # :128
    d_dns = sim.dist(num_bytes(expo(192.0 * B), header=68 * B, upper=576 * B))  # :171
    d_gap = sim.dist(expo(2.0 * MIN))                                         # :249
    d_delta = sim.dist(normal(0.0 * B, 10.0 * B))                             # :250
    d_name = sim.dist(distribution_index(num_workstations))                   # :371
    c_timeout = sim.const(5.0)                                                # :263

    # dhcp_serve, network_simple.py:152-167
    a.program("dhcp_serve", PROG_DHCP_SERVE)
    a.open(67)
    a.label("loop")
    a.recv()
    a.jeqi(PKT_TAG, TAG_DISCOVER, "offer")
    a.jeqi(PKT_TAG, TAG_REQUEST, "ack")
    a.jmp("loop")
    a.label("offer")
    a.drawi(R0, d_dhcp)
    a.send(DST_PKT_SRC, R0, TAG_OFFER)
    a.jmp("loop")
    a.label("ack")
    a.drawi(R0, d_dhcp)
    a.send(DST_PKT_SRC, R0, TAG_ACK)
    a.jmp("loop")

    # workstation_blinking + get_address_from_dhcp, network_simple.py:106-121,131-149
    a.program("blinking", PROG_BLINKING)
    a.label("loop")
    a.node_sleep()
    a.advance(d_sleeping)
    a.open(68)
    a.drawi(R0, d_dhcp)
    a.send(DST_BCAST, R0, TAG_DISCOVER, port=67)
    a.recv()
    a.drawi(R0, d_dhcp)
    a.send(DST_PKT_SRC, R0, TAG_REQUEST)
    a.recv()
    a.close()
    a.advance(d_ready)
    a.node_wake()
    a.advance(d_awake)
    a.jmp("loop")

    # mdns_daemon, network_simple.py:174-222
    a.program("mdns", PROG_MDNS)
    a.label("top")
    a.list_clear()
    a.open(5353)
    a.label("loop")
    if fused:
        # the three lines below plus the "not for me" exits of the dispatch, as one instruction
        a.recv_filter(R0, sim.packet_filter(always=[TAG_QUERY], if_host=[TAG_RESOLVE], if_listed=[TAG_IAM]), "asleep")
    else:
        a.wait_awake(R0)             # with ws.awake(): wait, remember when we awoke
        a.recv()
        a.jasleep(R0, "asleep")      # leaving the with: FellAsleep if the machine slept meanwhile
    a.jeqi(PKT_TAG, TAG_QUERY, "query")
    a.jeqi(PKT_TAG, TAG_IAM, "iam")
    a.jeqi(PKT_TAG, TAG_RESOLVE, "resolve")
    a.jmp("loop")
    a.label("query")
    a.list_push()
    a.send(DST_BCAST, PKT_SIZE, TAG_RESOLVE, port=5353, f0=PKT_F0)
    a.jmp("loop")
    a.label("iam")
    a.list_take(PKT_F0, "loop")
    a.send(DST_ENT, PKT_SIZE, TAG_IAM, f0=PKT_F0, f1=PKT_F1)
    a.jmp("loop")
    a.label("resolve")
    a.jne(PKT_F0, NODE_HOST, "loop")
    a.drawi(R1, d_dns)
    a.send(DST_BCAST, R1, TAG_IAM, port=5353, f0=NODE_HOST, f1=NODE_ADDR)
    a.jmp("loop")
    a.label("asleep")
    a.close()
    a.jmp("top")

    # llmnr_daemon, network_simple.py:225-246
    a.program("llmnr", PROG_LLMNR)
    a.label("top")
    a.open(5355)
    a.label("loop")
    if fused:
        a.recv_filter(R0, sim.packet_filter(if_host=[TAG_RESOLVE]), "asleep")
    else:
        a.wait_awake(R0)
        a.recv()
        a.jasleep(R0, "asleep")
    a.jnei(PKT_TAG, TAG_RESOLVE, "loop")
    a.jne(PKT_F0, NODE_HOST, "loop")
    a.drawi(R1, d_dns)
    a.send(DST_PKT_SRC, R1, TAG_IAM, f0=NODE_HOST, f1=NODE_ADDR)
    a.jmp("loop")
    a.label("asleep")
    a.close()
    a.jmp("top")

    # client_activity, network_simple.py:283-305
    a.program("client", PROG_CLIENT)
    a.label("loop")
    a.wait_awake()                   # ws.wait_until_awake(logger)
    a.wait_awake(R0)                 # with ws.awake():
    a.advance(d_gap)
    a.jasleep(R0, "loop")            # FellAsleep -> caught -> next round
    a.label("pick")
    a.drawi(R1, d_name)              # next(name_next_query): index into names_ws
    a.addi(R1, R1, 1)                # hostname id of "Workstation-{index+1}"
    a.jeq(R1, NODE_HOST, "pick")
    a.stat(STAT_QUERIES)
    a.drawi(R0, d_dns)               # size_packet_base
    a.spawn("query_mdns")
    a.spawn("query_llmnr")
    a.jmp("loop")

    # _query / _query_mdns / _query_llmnr, network_simple.py:253-280 with the timeout guard of :84-98
    for name, prog in (("query_mdns", PROG_QUERY_MDNS), ("query_llmnr", PROG_QUERY_LLMNR)):
        a.program(name, prog)
        a.open(0)
        a.draw_addi(R0, R0, d_delta)                 # int(size_packet_base + next(size_packet_delta))
        if name == "query_mdns":
            a.send(DST_SELF, R0, TAG_QUERY, port=5353, f0=R1)
        else:
            a.send(DST_BCAST, R0, TAG_RESOLVE, port=5355, f0=R1)
        a.wait_awake(R2)                             # with ws.awake(),
        a.mov(R3, SELF)
        a.spawn("timeout_helper", rd=R3)             #      ws.timeout(5.0): helper gets our handle, we keep its
        a.recv(handler="timed_out")
        a.stat(STAT_RESOLVED)
        a.throw(R3, EXC_COMPLETE)                    # leaving timeout(): stop the helper
        a.jasleep(R2, "end")                         # leaving awake(): FellAsleep, swallowed at :268
        a.label("end")
        a.close()
        a.exit()
        a.label("timed_out")                         # _Complete in the body -> Workstation.Timeout, swallowed :266
        a.stat(STAT_TIMEOUTS)
        a.close()
        a.exit()

    # Workstation._wait_for_timeout, network_simple.py:92-98
    a.program("timeout_helper", PROG_TIMEOUT_HELPER)
    a.advance_const(c_timeout, handler="done")
    a.throw(R3, EXC_COMPLETE)
    a.label("done")
    a.exit()


def build(cidr: str = "192.168.4.0/24", num_endpoints: Optional[int] = None, trace_records: int = 0,
          fused: bool = True, connection_records: int = 0, connection_window: int = 8) -> Simulator:
    """``init()`` of network_simple.py:311-387 (argument defaults :315-316).  ``fused=False`` assembles the daemon
    loops from the plain wait/recv/branch instructions instead of the fused loop head (same behaviour, used by the
    tests to check one encoding against the other).  ``connection_records`` > 0 keeps the per-socket packet
    correlation of overrides.py:102-186 and logs its connection records (capacity per replica); the receive loops are
    then assembled from plain instructions, since every received packet has to be correlated."""
    sim = Simulator()
    if connection_records:
        sim.asm.log_connections = True
        sim.caps.update(connection_records=connection_records, connection_window=connection_window)
    num_addresses = ip_network(cidr).num_addresses - NUM_ADDRESSES_RESERVED
    if num_addresses < MIN_NUM_ENDPOINTS:
        raise ValueError(f"Unsuitable CIDR prefix for simulating a non-trivial network: {cidr}")
    net_local = Network(sim, cidr=cidr, bandwidth=constant(1 * GbPS), latency=normal(5 * MS, 1 * MS))
    router = Endpoint("Router", net_local)
    workstations = []
    if num_endpoints is None:
        num_endpoints = max(MIN_NUM_ENDPOINTS, int(0.2 * num_addresses))
    num_endpoints = max(MIN_NUM_ENDPOINTS, num_endpoints)
    num_workstations = num_endpoints - 1
    num_mdns = int(0.9 * num_workstations)

    _programs(sim, num_workstations, fused)

    dhcp_server = Endpoint("DHCPServer", net_local)
    dhcp_server.install("dhcp_serve")
    for n in range(num_workstations):
        ws = Workstation(f"Workstation-{n + 1}", net_local, host_id=n + 1)
        ws.install("blinking")
        ws.install("client")
        ws.install("mdns" if n <= num_mdns else "llmnr")   # sic, network_simple.py:380
        workstations.append(ws)
    sim.router, sim.workstations = router, workstations
    sim.caps["trace_records"] = trace_records
    # per-replica shared-memory budget (every structure continues in the HBM arena when it fills, except the
    # process and socket tables): 148 resident processes + 72 transient ones (18 query rounds in flight at once);
    # sized so that 12 replicas (warps) fit one SM's shared memory
    sim.caps.update(max_events=128, max_procs=220, max_sockets=96)
    sim.address_tags = 1 << TAG_IAM      # the "iam" payloads carry PayloadDictionaryType.ADDRESS (network_simple.py:215,244)
    sim.num_workstations = num_workstations
    return sim


def build_malware(trace_records: int = 0, fused: bool = True, connection_records: int = 0,
                  connection_window: int = 8, beacon_gap: Optional[Dist] = None,
                  privileged_gap: Optional[Dist] = None, c2_targets: Optional[Sequence[Tuple[int, int]]] = None,
                  beacon_size: Optional[Dist] = None, infected_index: Optional[int] = None) -> Simulator:
    """network_simple + the beaconing-malware overlay of ``examples/network_simple/malware_simple.py:53-125``
    (BASELINE.json config 3).  One workstation drawn per replica at set-up runs ``call_home`` (beacon every expo(10) s)
    and ``no_good`` (every expo(100) s); the router's ``route`` relays the 100-byte beacons over the ``Internet``
    network (overrides.py:443-456) to the command-and-control host 1.2.3.4:666.  The overlay's own statistics are the
    user counters STAT_BEACONS / STAT_ROUTED / STAT_C2_RECEIVED.

    The keyword parameters are those of ``examples/backdoor/backdoor.py`` (which itself cannot run: it calls
    ``tcp_connect`` and ``random_network_activity``, absent from the reference): ``beacon_gap`` (:34, uniform(2.1 h,
    4.9 h)), ``c2_targets`` (:11-20, a beacon goes to ``next(distribution(targets))``), ``beacon_size`` (:21),
    ``infected_index`` (:69, workstations[23]); see ``build_backdoor``."""
    sim = build(trace_records=trace_records, fused=fused, connection_records=connection_records,
                connection_window=connection_window)
    a = sim.asm
    www = Network(sim, "0.0.0.0/0", latency=bounded(expo(1 * S), lower=20 * MS), bandwidth=expo(10 * MbPS))
    d_beacon = sim.dist(beacon_gap if beacon_gap is not None else expo(10))
    d_privileged = sim.dist(privileged_gap if privileged_gap is not None else expo(100))
    router_wan = sim.router.link_to(www)                          # malware_simple.py:117
    targets = list(c2_targets) if c2_targets else [(BAD_NEWS_BEAR_ADDR, BAD_NEWS_BEAR_PORT)]
    if len(targets) > 16:
        raise ValueError("at most 16 command-and-control targets")
    bears = [Endpoint("Mama Bear" if k == 0 else f"Mama Bear {k + 1}", www, addr) for k, (addr, _) in enumerate(targets)]
    bear = bears[0]

    a.program("call_home", PROG_CALL_HOME, tags=[Tag.MALWARE])    # malware_simple.py:53-58 (@malware)
    a.label("loop")
    a.advance(d_beacon)
    a.spawn("exfiltrate")
    a.jmp("loop")

    a.program("no_good", PROG_NO_GOOD, tags=[Tag.MALWARE])        # :61-65 (@malware)
    a.label("loop")
    a.advance(d_privileged)
    a.spawn("sudo_transmit")
    a.jmp("loop")

    a.program("sudo_transmit", PROG_SUDO_TRANSMIT, tags=[Tag.VULNERABLE])   # :68-72 (@tagged(Tag.VULNERABLE))
    a.spawn("exfiltrate")
    a.exit()

    a.program("exfiltrate", PROG_EXFILTRATE)                      # :75-82
    a.stat(STAT_BEACONS)
    a.stat_first(STAT_FIRST_BEACON_US)                            # time to the first beacon (a config-3 statistic)
    if c2_targets:                                                # backdoor.py:36 next(location_c2)
        a.drawi(R0, sim.dist(distribution_index(len(targets))))
        for k in range(1, len(targets)):
            a.jeqi(R0, k, f"target{k}")
    for k, (addr, port) in enumerate(targets):
        if k:
            a.label(f"target{k}")
        a.li(R1, port)
        a.gset(1, R1)                                             # payload["dest_addr"] = Location(addr, port)
        a.li(R1, addr)
        if k + 1 < len(targets):
            a.jmp("chosen")
    a.label("chosen")
    if beacon_size is not None:                                   # backdoor.py:37 next(len_beacon)
        a.drawi(R0, sim.dist(beacon_size))
    else:
        a.li(R0, 100)
    a.open(BAD_NEWS_BEAR_PORT)
    a.li(R2, sim.router.address)
    a.li(R3, LOCAL_ROUTER_PORT)
    a.send(DST_REGS, R0, TAG_EXFIL, f0=R1, f1=G0 + 1)
    a.close()
    a.exit()

    a.program("route", PROG_ROUTE)                                # :96-107 (the WAN-side socket only sends)
    a.open(LOCAL_ROUTER_PORT)
    a.open_at(router_wan, LOCAL_ROUTER_PORT)
    a.label("loop")
    a.recv()
    a.stat(STAT_ROUTED)
    a.stat_add(STAT_EXFIL_BYTES, PKT_SIZE)                        # bytes that left the local network
    a.mov(R2, _PKT_F0)
    a.mov(R3, _PKT_F1)
    a.forward(router_wan, LOCAL_ROUTER_PORT)
    a.jmp("loop")

    for port in sorted({port for _, port in targets}):
        a.program(f"command_and_control_{port}", PROG_C2, tags=[Tag.MALWARE])   # :85-92 (@malware)
        a.open(port)
        a.label("loop")
        a.recv()
        a.stat(STAT_C2_RECEIVED)
        a.jmp("loop")

    if infected_index is None:
        infected = DrawnNode(sim, sim.workstations)               # :114 next(distribution(workstation_list))
    else:
        infected = sim.workstations[infected_index]               # backdoor.py:69
    infected.install("call_home")
    infected.install("no_good")
    sim.router.install("route")
    for b, (_, port) in zip(bears, targets):
        b.install(f"command_and_control_{port}")
    sim.router_wan, sim.bear, sim.bears = router_wan, bear, bears
    sim.caps.update(max_procs=224, max_sockets=100 + len(targets))
    return sim


# examples/backdoor/backdoor.py:11-22,34,69
BACKDOOR_C2 = [(int(ip_address("2.7.45.67")), 80), (int(ip_address("198.51.100.7")), 8080),      # "rom.com"
               (int(ip_address("8.45.90.222")), 80), (int(ip_address("198.51.100.9")), 8081)]    # "intelligent.design"


def build_backdoor(**kw) -> Simulator:
    """The malware overlay with the parameters of ``examples/backdoor/backdoor.py``: 50 workstations, workstation 24
    (index 23) infected, a beacon every uniform(2.1 h, 4.9 h) to one of four C2 locations, beacon length
    ``num_bytes(expo(1024 B), header=128 B, upper=20 MB)``; meant to run 72 h (:70).  The two locations the example
    names by host name get documentation addresses (the reference has no resolver to give them one)."""
    return build_malware(beacon_gap=uniform(2.1 * H, 4.9 * H), c2_targets=BACKDOOR_C2,
                         beacon_size=num_bytes(expo(1024 * B), header=128 * B, upper=20 * MB), infected_index=23, **kw)
