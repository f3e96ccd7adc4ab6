"""
ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's flagship workload.

PARITY PINNED: ``oracle/ref_example.py`` executes the reference's own example files unmodified on the same Philox stream,
and ``tests/test_oracle.py`` holds this module's traces, connection records and draw counts against theirs, bit for bit.

``examples/network_simple`` cannot run against the shipped ``itsim`` package (it imports
``Payload``/``PayloadDictionaryType`` from ``itsim.network.packet`` and ``Node``/``Socket`` from ``itsim.machine``,
none of which exist; ``network_simple_overrides.py:1-6`` calls itself deprecated).  This module restates, on the
greensim shim, exactly the behaviour those two files specify:

* flat network + override node/socket classes  <- ``examples/network_simple/network_simple_overrides.py:65-479``
* workstation wake/sleep, DHCP toy exchange, mDNS / LLMNR responders, client queries with a 5 s timeout
                                                <- ``examples/network_simple/network_simple.py:42-387``

Recorded deviations (all forced by F4 / F9, none changes event order):

* payloads are plain dicts keyed ``"content"``, ``"hostname"``, ``"address"`` (the ``Payload`` classes are missing);
* ``Node._as_source_bind`` is called at ``network_simple_overrides.py:338`` but defined nowhere: restated as
  "None -> (default address, port 0); an int -> (default address, that port)";
* log lines are dropped (they only format strings); the per-socket connection telemetry of
  ``network_simple_overrides.py:102-120,152-186`` is kept and collected in ``Model.connections``;
* every variate generator draws from the one per-replica Philox stream (``oracle/philox.py``), in execution
  order, instead of the process-global Mersenne Twister;
* processes are labelled (program id, node index) for the tracer -- pure observation.
"""
import os
import sys
from collections import OrderedDict
from contextlib import contextmanager
from ipaddress import ip_network
from itertools import cycle
from queue import Queue as _Fifo
from typing import Any, Dict, Iterator, List, Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import greensim  # noqa: E402  (the shim in oracle/greensim)
from greensim import Process, Signal, Simulator, add, advance, now, tagged  # noqa: E402
from greensim.tags import Tags  # noqa: E402
from greensim.random import bounded, constant, distribution, expo, linear, normal, project_int  # noqa: E402
import greensim.random as gsrandom  # noqa: E402

from philox import PhiloxRandom  # noqa: E402
import evtrace as tr  # noqa: E402

assert getattr(greensim, "GREENSIM_TAG_ATTRIBUTE", None), "oracle/greensim shim expected"

# ---- itsim/units.py:2-14, quirks included (US = MS*1e-6 = 1e-9 s; bandwidth units are bytes/s) -------------
S = 1.0
MIN = 60 * S
H = 60 * MIN
MS = S * 1.0e-3
US = MS * 1.0e-6
KbPS = 1024 / 8
MbPS = 1024 * KbPS
GbPS = 1024 * MbPS

# ---- program ids shared with the engine (itsim_b200/netsimple.py) -----------------------------------------
PROG_DHCP_SERVE, PROG_BLINKING, PROG_CLIENT, PROG_MDNS, PROG_LLMNR = 1, 2, 3, 4, 5
PROG_QUERY_MDNS, PROG_QUERY_LLMNR, PROG_TIMEOUT_HELPER, PROG_TX, PROG_DELIVER = 6, 7, 8, 9, 10
PROG_CALL_HOME, PROG_NO_GOOD, PROG_SUDO_TRANSMIT, PROG_EXFILTRATE, PROG_ROUTE, PROG_C2 = 11, 12, 13, 14, 15, 16
EXC_COMPLETE = 1

Location = Tuple[int, int]  # (IPv4 address as int, port); hashable like itsim.network.location.Location


def num_bytes(gen: Iterator[float], header: int = 0, upper: Optional[float] = None) -> Iterator[int]:
    """itsim/random.py:14-29."""
    yield from project_int(bounded(linear(gen, 1.0, header), 0.0, upper))


class Packet:
    """itsim/network/packet.py:9-60 (value object; ``len`` = byte size)."""
    __slots__ = ("source", "dest", "byte_size", "payload")

    def __init__(self, source: Location, dest: Location, byte_size: int, payload: Dict[str, Any]) -> None:
        self.source = source
        self.dest = dest
        self.byte_size = byte_size
        self.payload = payload

    def __len__(self) -> int:
        return self.byte_size


class Network:
    """network_simple_overrides.py:189-297 -- one broadcast domain; every transmission is its own process."""

    def __init__(self, model: "Model", cidr: str, latency: Iterator[float], bandwidth: Iterator[float]) -> None:
        self.model = model
        self.sim = model.sim
        self.cidr = ip_network(cidr)
        self._latency = bounded(latency, lower=0)            # overrides.py:202
        self._bandwidth = bounded(bandwidth, lower=0.125)    # overrides.py:203
        self._nodes: "OrderedDict[int, Node]" = OrderedDict()
        self._forwarders: "OrderedDict[Any, Node]" = OrderedDict()
        self._top_free = (int(a) for a in self.cidr.hosts())  # overrides.py:207-213 (no skipped addresses)
        self.address_broadcast = int(self.cidr.broadcast_address)

    def link(self, node: "Node", address: Optional[int] = None) -> int:
        if address is None:
            address = next(self._top_free)
        if address in self._nodes:
            raise ValueError("address in use")
        self._nodes[address] = node
        return address

    def _address_of(self, node: "Node") -> int:
        for address, member in self._nodes.items():
            if member is node:
                return address
        return -1

    def transmit(self, packet: Packet) -> None:
        """overrides.py:273-291."""
        dest_host = packet.dest[0]
        if dest_host == self.address_broadcast:
            receivers = self._nodes.values()
        elif dest_host in self._nodes:
            receivers = [self._nodes[dest_host]]
        else:
            raise LookupError(f"cannot forward to {dest_host}")  # CannotForward, overrides.py:293-297

        def transmission() -> None:
            advance(next(self._latency) + len(packet) / next(self._bandwidth))
            for node in receivers:
                index = self.model.node_of_address.get(self._address_of(node), node).index
                proc = self.sim.add(node._receive, packet)
                tr.Tracer.label(proc, tr.ROLE_DELIVER, PROG_DELIVER, index)

        proc = self.sim.add(transmission)
        tr.Tracer.label(proc, tr.ROLE_TX, PROG_TX, self.model.node_of_address[packet.source[0]].index,
                        packet.byte_size)
        self.model.stats["packets_sent"] += 1
        self.model.stats["bytes_sent"] += packet.byte_size


class _NetworkLink:
    """overrides.py:462-479 -- per (node, address) port table with its persistent free-port cursor."""

    def __init__(self, address: int, network: Network) -> None:
        self.address = address
        self.network = network
        self.ports: "OrderedDict[int, Any]" = OrderedDict()
        self._seq_ports_unprivileged = cycle(range(1024, 65536))

    def get_port_free(self) -> int:
        if len(self.ports) >= 60000:
            raise RuntimeError("too many ports")
        for port in self._seq_ports_unprivileged:
            if port not in self.ports:
                return port
        raise AssertionError


class Tag(Tags):
    """itsim/__init__.py:9-11."""
    MALWARE = 0
    VULNERABLE = 1


def current_tags() -> int:
    """Bit mask (bit = tag value) of ``Process.current().iter_tags()``: what malware_simple.py:46 prints per log line."""
    mask = 0
    for t in Process.current().iter_tags():
        mask |= 1 << t.value
    return mask


class Socket:
    """overrides.py:65-186."""

    def __init__(self, src: Location, node: "Node") -> None:
        self._src = src
        self._node = node
        self._packet_queue: "_Fifo[Packet]" = _Fifo()
        self._packet_signal = Signal()
        self._packet_signal.turn_off()
        self._outbound: "OrderedDict[Location, Tuple[Packet, float]]" = OrderedDict()
        self._inbound: "OrderedDict[Location, Tuple[Packet, float]]" = OrderedDict()

    # -- telemetry correlation (overrides.py:102-120) ------------------------------------------------
    def _correlate_outbound(self, dest: Location, packet: Packet) -> None:
        if dest in self._inbound:
            self._log_pair(self._inbound[dest][0], packet, self._inbound[dest][1], True)
            del self._inbound[dest]
        else:
            if dest in self._outbound:
                self._log_pair(None, self._outbound[dest][0], self._outbound[dest][1], False)
            self._outbound[dest] = (packet, now())
            self._forget_oldest(self._outbound)

    def _correlate_inbound(self, src: Location, packet: Packet) -> None:
        if src in self._outbound:
            self._log_pair(packet, self._outbound[src][0], self._outbound[src][1], False)
            del self._outbound[src]
        else:
            self._inbound[src] = (packet, now())
            self._forget_oldest(self._inbound)

    def _forget_oldest(self, filed: "OrderedDict[Location, Tuple[Packet, float]]") -> None:
        """Not in the reference, whose dictionaries never forget (a daemon socket's inbound dictionary grows by one
        entry per broadcast it hears until the workstation falls asleep and the socket closes).  ``Model.conn_window``
        bounds each dictionary to its newest entries: the oldest one is logged the way ``flush_logs`` would log it at
        close time, just earlier.  None keeps the reference's behaviour; a window the dictionaries never reach gives
        the reference's records exactly (tests/test_connections.py)."""
        window = self._node.model.conn_window
        if window is not None and len(filed) > window:
            _, (packet, start) = filed.popitem(last=False)
            if filed is self._outbound:
                self._log_pair(None, packet, start, False)
            else:
                self._log_pair(packet, None, start, True)

    def _log_pair(self, packet_in: Optional[Packet], packet_out: Optional[Packet], start: float,
                  is_inbound: bool) -> None:
        """overrides.py:152-177: one connection record (field semantics kept; JSON formatting dropped)."""
        remote_ip = 0
        if packet_out is not None:
            remote_ip = packet_out.payload["address"] if "address" in packet_out.payload else packet_out.dest[0]
        self._node.model.connections.append((
            now(), self._node.index,
            0 if packet_in is None else packet_in.dest[0], 0 if packet_in is None else packet_in.dest[1],
            1 if is_inbound else 0, remote_ip, 0 if packet_out is None else packet_out.dest[1],
            0 if packet_out is None else packet_out.byte_size, 0 if packet_in is None else packet_in.byte_size,
            start))

    def flush_logs(self) -> None:
        for packet, start in self._outbound.values():
            self._log_pair(None, packet, start, False)
        for packet, start in self._inbound.values():
            self._log_pair(packet, None, start, True)

    # -- data path (overrides.py:122-147) ------------------------------------------------------------
    def send(self, dest: Location, byte_size: int, payload: Dict[str, Any]) -> None:
        packet = Packet(self._src, dest, byte_size, payload)
        self._node._send_to_network(packet)
        self._correlate_outbound(dest, packet)

    def broadcast(self, port: int, byte_size: int, payload: Dict[str, Any]) -> None:
        self.send((self._node._broadcast_address(self._src[0]), port), byte_size, payload)

    def _enqueue(self, packet: Packet) -> None:
        self._packet_queue.put(packet)
        self._packet_signal.turn_on()

    def recv(self) -> Packet:
        while self._packet_queue.empty():
            self._packet_signal.wait()
        packet = self._packet_queue.get()
        self._correlate_inbound(packet.source, packet)
        if self._packet_queue.empty():
            self._packet_signal.turn_off()
        self._node.model.stats["packets_received"] += 1
        self._node.model.stats["bytes_received"] += packet.byte_size
        return packet


class Node:
    """overrides.py:300-388 plus Endpoint (overrides.py:403-433)."""

    def __init__(self, model: "Model", name: str, network: Network, address: Optional[int] = None) -> None:
        self.model = model
        self.name = name
        self.sim = model.sim
        self.index = len(model.nodes)
        model.nodes.append(self)
        self._networks: "OrderedDict[int, _NetworkLink]" = OrderedDict()
        self._sockets: "OrderedDict[Location, Socket]" = OrderedDict()
        address = network.link(self, address)
        self._networks[address] = _NetworkLink(address, network)
        model.node_of_address[address] = self

    def link_to(self, network: Network, address: Optional[int] = None) -> int:
        """overrides.py:307-312: a further address on another network.  The trace numbers that attachment like a
        node of its own (the engine models it as one), hence the placeholder in ``model.nodes``."""
        address = network.link(self, address)
        self._networks[address] = _NetworkLink(address, network)
        alias = _Attachment(self, len(self.model.nodes))
        self.model.nodes.append(alias)
        self.model.node_of_address[address] = alias
        return address

    @property
    def address_default(self) -> int:
        return next(iter(self._networks.keys()))

    def _broadcast_address(self, src_addr: int) -> int:
        return self._networks[src_addr].network.address_broadcast

    @contextmanager
    def bind(self, port: Any = None) -> Iterator[Location]:
        src: Location = port if isinstance(port, tuple) else (self.address_default, port or 0)
        link = self._networks[src[0]]
        if src[1] == 0:
            src = (src[0], link.get_port_free())
        if src[1] in link.ports:
            raise RuntimeError("port already in use")
        link.ports[src[1]] = Process.current()
        try:
            yield src
        finally:
            del link.ports[src[1]]

    @contextmanager
    def open_socket(self, port: Any = None) -> Iterator[Socket]:
        with self.bind(port) as src:
            sock = Socket(src, self)
            alias = (self._broadcast_address(src[0]), src[1])  # also listens on the broadcast address
            self._sockets[src] = sock
            self._sockets[alias] = sock
            self.model.net_events.append((now(), self.index, src[1], 1, current_tags()))   # network_event "open"
            try:
                yield sock
            finally:
                self.model.net_events.append((now(), self.index, src[1], 2, current_tags()))   # "close"
                sock.flush_logs()
                del self._sockets[src]
                del self._sockets[alias]

    def _send_to_network(self, packet: Packet) -> None:
        src = packet.source
        if src[0] not in self._networks or src[1] not in self._networks[src[0]].ports:
            raise RuntimeError("no network linked")
        self._networks[src[0]].network.transmit(packet)

    def _receive(self, packet: Packet) -> None:
        if packet.dest in self._sockets:
            self._sockets[packet.dest]._enqueue(packet)

    def install(self, prog: int, fn_software: Any, *args: Any) -> None:
        proc = self.sim.add(fn_software, self, *args)
        tr.Tracer.label(proc, tr.ROLE_MODEL, prog, self.index)


class _Attachment:
    """A node's attachment to a further network, numbered like a node (see ``Node.link_to``)."""

    def __init__(self, node: Node, index: int) -> None:
        self.node, self.index = node, index

    def _receive(self, packet: Packet) -> None:
        self.node._receive(packet)


class FellAsleep(Exception):
    pass


class _Complete(Exception):
    pass


class QueryTimeout(Exception):
    pass


class Workstation(Node):
    """network_simple.py:42-98."""

    def __init__(self, model: "Model", name: str, network: Network) -> None:
        super().__init__(model, name, network)
        self._wake = Signal().turn_off()
        self._time_awoken = -1.0

    def wake_up(self) -> None:
        self._wake.turn_on()
        self._time_awoken = now()

    def sleep(self) -> None:
        self._wake.turn_off()
        self._time_awoken = -1.0

    def wait_until_awake(self) -> None:
        self._wake.wait()

    @contextmanager
    def awake(self) -> Iterator["Workstation"]:
        self.wait_until_awake()
        time_at_start = self._time_awoken
        yield self
        if self._time_awoken != time_at_start:
            raise FellAsleep()

    @contextmanager
    def timeout(self, delay: float) -> Iterator["Workstation"]:
        helper = add(self._wait_for_timeout, delay, Process.current())
        tr.Tracer.label(helper, tr.ROLE_MODEL, PROG_TIMEOUT_HELPER, self.index)
        try:
            yield self
            _throw_process(helper, _Complete())
        except _Complete:
            raise QueryTimeout()

    def _wait_for_timeout(self, delay: float, process_main: Process) -> None:
        try:
            advance(delay)
            _throw_process(process_main, _Complete())
        except _Complete:
            pass


def _throw_process(proc: Process, exc: BaseException) -> None:
    """network_simple.py:37-38."""
    proc.rsim()._schedule(0.0, proc.throw, exc)


class Model:
    """One replica of network_simple: ``init()`` of network_simple.py:311-387 with its module-level generators."""

    def __init__(self, seed: int = 0, replica: int = 0, cidr: str = "192.168.4.0/24",
                 num_endpoints: Optional[int] = None, trace: bool = True, trace_limit: Optional[int] = None,
                 conn_window: Optional[int] = None) -> None:
        self.conn_window = conn_window
        self.rng = PhiloxRandom(seed, replica)
        self.sim = Simulator()
        self.nodes: List[Node] = []
        self.node_of_address: Dict[int, Node] = {}
        self.connections: List[tuple] = []
        self.net_events: List[tuple] = []        # (t, node, port, 1 = open | 2 = close): itsim/schemas/items.py:83-131
        self.stats = {"packets_sent": 0, "bytes_sent": 0, "packets_received": 0, "bytes_received": 0,
                      "queries": 0, "resolved": 0, "timeouts": 0}
        self.tracer = tr.Tracer(self.sim, keep_records=trace, limit=trace_limit)
        self.tracer.exc_code = lambda exc: EXC_COMPLETE if isinstance(exc, _Complete) else 0
        self.tracer.deliver_aux = self._deliver_aux

        # module-level generators of network_simple.py:101-103,128,171,249-250
        self.delay_awake = expo(2.0 * H)
        self.delay_sleeping = expo(10.0 * MIN)
        self.delay_setup_to_ready = expo(5.0 * US)
        self.size_packet_dhcp = num_bytes(normal(100.0, 30.0), header=240)
        self.size_packet_dns = num_bytes(expo(192.0), header=68, upper=576)
        self.delay_identity_queries = expo(2.0 * MIN)
        self.size_packet_delta = normal(0.0, 10.0)

        num_addresses = ip_network(cidr).num_addresses - 2
        self.network = Network(self, cidr, latency=normal(5 * MS, 1 * MS), bandwidth=constant(1 * GbPS))
        self.router = Node(self, "Router", self.network)
        if num_endpoints is None:
            num_endpoints = max(3, int(0.2 * num_addresses))
        num_endpoints = max(3, num_endpoints)
        num_workstations = num_endpoints - 1
        num_mdns = int(0.9 * num_workstations)
        self.num_workstations, self.num_mdns = num_workstations, num_mdns

        self.dhcp_server = Node(self, "DHCPServer", self.network)
        self.dhcp_server.install(PROG_DHCP_SERVE, self.dhcp_serve)

        self.names_ws: List[str] = []
        self.name_next_query = distribution(self.names_ws)
        self.workstations: List[Workstation] = []
        for n in range(num_workstations):
            name = f"Workstation-{n + 1}"
            self.names_ws.append(name)
            ws = Workstation(self, name, self.network)
            ws.install(PROG_BLINKING, self.workstation_blinking)
            ws.install(PROG_CLIENT, self.client_activity)
            if n <= num_mdns:  # sic: network_simple.py:380 gives mDNS to num_mdns + 1 workstations
                ws.install(PROG_MDNS, self.mdns_daemon)
            else:
                ws.install(PROG_LLMNR, self.llmnr_daemon)
            self.workstations.append(ws)

    # -- observation -----------------------------------------------------------------------------------
    @staticmethod
    def _deliver_aux(proc: Any, ev: Any) -> int:
        node, packet = proc._body.__self__, ev.args[0]
        accepted = 1 if packet.dest in node._sockets else 0
        return (packet.byte_size & 0x7FFFFFFF) | (accepted << 31)

    def run(self, duration: float) -> None:
        gsrandom.set_default_rng(self.rng)
        try:
            self.sim.run(duration)
        finally:
            gsrandom.set_default_rng(None)

    def close(self) -> None:
        self.sim._clear()

    # -- process bodies ----------------------------------------------------------------------------------
    def workstation_blinking(self, ws: Workstation) -> None:
        """network_simple.py:106-121."""
        while True:
            ws.sleep()
            advance(next(self.delay_sleeping))
            self.get_address_from_dhcp(ws)
            advance(next(self.delay_setup_to_ready))
            ws.wake_up()
            advance(next(self.delay_awake))

    def get_address_from_dhcp(self, ws: Workstation) -> None:
        """network_simple.py:131-149."""
        with ws.open_socket(68) as socket:
            socket.broadcast(67, next(self.size_packet_dhcp), {"content": "DHCPDISCOVER"})
            packet_offer = socket.recv()
            socket.send(packet_offer.source, next(self.size_packet_dhcp), {"content": "DHCPREQUEST"})
            socket.recv()

    def dhcp_serve(self, node: Node) -> None:
        """network_simple.py:152-167."""
        responses = {"DHCPDISCOVER": "DHCPOFFER", "DHCPREQUEST": "DHCPACK"}
        with node.open_socket(67) as socket:
            while True:
                packet_client = socket.recv()
                type_msg = packet_client.payload["content"]
                if type_msg in responses:
                    socket.send(packet_client.source, next(self.size_packet_dhcp), {"content": responses[type_msg]})

    def mdns_daemon(self, ws: Workstation) -> None:
        """network_simple.py:174-222."""
        while True:
            queries: List[Packet] = []
            try:
                with ws.open_socket(5353) as socket:
                    while True:
                        with ws.awake():
                            packet = socket.recv()
                        entries = packet.payload
                        content_resolve = entries["content"] == "resolve"
                        hostname_match = entries["hostname"] == ws.name
                        if entries["content"] == "query":
                            queries.append(packet)
                            socket.broadcast(5353, len(packet),
                                             {"content": "resolve", "hostname": entries["hostname"]})
                        elif entries["content"] == "iam":
                            i_del = None
                            for i, packet_query in enumerate(queries):
                                if packet_query.payload["hostname"] == entries["hostname"]:
                                    i_del = i
                                    socket.send(packet_query.source, len(packet), packet.payload)
                                    break
                            if i_del is not None:
                                del queries[i_del]
                        elif content_resolve and hostname_match:
                            socket.broadcast(5353, next(self.size_packet_dns),
                                             {"content": "iam", "hostname": ws.name, "address": ws.address_default})
            except FellAsleep:
                pass

    def llmnr_daemon(self, ws: Workstation) -> None:
        """network_simple.py:225-246."""
        while True:
            try:
                with ws.open_socket(5355) as socket:
                    while True:
                        with ws.awake():
                            packet = socket.recv()
                        entries = packet.payload
                        if entries["content"] == "resolve" and entries["hostname"] == ws.name:
                            socket.send(packet.source, next(self.size_packet_dns),
                                        {"content": "iam", "hostname": ws.name, "address": ws.address_default})
            except FellAsleep:
                pass

    @contextmanager
    def _query(self, ws: Workstation, size_packet_base: int) -> Iterator[Tuple[Socket, int]]:
        """network_simple.py:253-270."""
        try:
            with ws.open_socket() as socket:
                yield (socket, int(size_packet_base + next(self.size_packet_delta)))
                try:
                    with ws.awake(), ws.timeout(5.0):
                        socket.recv()
                        self.stats["resolved"] += 1
                except QueryTimeout:
                    self.stats["timeouts"] += 1
        except FellAsleep:
            pass

    def _query_mdns(self, ws: Workstation, size_packet_base: int, payload: Dict[str, Any]) -> None:
        """network_simple.py:273-275."""
        with self._query(ws, size_packet_base) as (socket, size_packet):
            socket.send((ws.address_default, 5353), size_packet, payload)

    def _query_llmnr(self, ws: Workstation, size_packet_base: int, payload: Dict[str, Any]) -> None:
        """network_simple.py:278-280."""
        with self._query(ws, size_packet_base) as (socket, size_packet):
            socket.broadcast(5355, size_packet, payload)

    def client_activity(self, ws: Workstation) -> None:
        """network_simple.py:283-305."""
        while True:
            ws.wait_until_awake()
            try:
                with ws.awake():
                    advance(next(self.delay_identity_queries))
                while True:
                    target_query = next(self.name_next_query)
                    if target_query != ws.name:
                        break
                self.stats["queries"] += 1
                size_packet_base = next(self.size_packet_dns)
                proc = add(self._query_mdns, ws, size_packet_base, {"content": "query", "hostname": target_query})
                tr.Tracer.label(proc, tr.ROLE_MODEL, PROG_QUERY_MDNS, ws.index)
                proc = add(self._query_llmnr, ws, size_packet_base, {"content": "resolve", "hostname": target_query})
                tr.Tracer.label(proc, tr.ROLE_MODEL, PROG_QUERY_LLMNR, ws.index)
            except FellAsleep:
                pass


class MalwareModel(Model):
    """network_simple + the beaconing-malware overlay of examples/network_simple/malware_simple.py:53-125 (config C3).

    One workstation, chosen with ``next(distribution(workstation_list))`` at set-up (malware_simple.py:114: the
    replica's first draw), runs ``call_home`` (beacon every expo(10) s) and ``no_good`` (privileged read every
    expo(100) s); both end in ``exfiltrate``: 100 B to the router's port 777.  The router's ``route`` process relays
    every such packet over the ``Internet`` network (latency max(expo(1 s), 20 ms), bandwidth max(expo(10 MbPS),
    0.125): overrides.py:443-456) to the command-and-control host 1.2.3.4:666."""

    BAD_NEWS_BEAR_ADDR = (1 << 24) | (2 << 16) | (3 << 8) | 4
    BAD_NEWS_BEAR_PORT = 666
    LOCAL_ROUTER_PORT = 777

    def __init__(self, seed: int = 0, replica: int = 0, trace: bool = True, trace_limit: Optional[int] = None,
                 conn_window: Optional[int] = None, beacon_gap: Any = None, privileged_gap: Any = None,
                 c2_targets: Optional[List[Location]] = None, beacon_size: Any = None,
                 infected_index: Optional[int] = None) -> None:
        """The keyword parameters are those of examples/backdoor/backdoor.py (itself not runnable: it calls
        ``tcp_connect`` and ``random_network_activity``, which do not exist): ``beacon_gap`` :34 uniform(2.1 h, 4.9 h),
        ``c2_targets`` :11-20 the C2 locations a beacon picks from with ``next(distribution(...))``, ``beacon_size``
        :21 num_bytes(expo(1024 B), header=128 B, upper=20 MB), ``infected_index`` :69 workstations[23].  Left at None
        they give malware_simple.py as written."""
        super().__init__(seed, replica, trace=trace, trace_limit=trace_limit, conn_window=conn_window)
        self.beacon_gap, self.privileged_gap, self.beacon_size = beacon_gap, privileged_gap, beacon_size
        self.c2_targets = c2_targets or [(self.BAD_NEWS_BEAR_ADDR, self.BAD_NEWS_BEAR_PORT)]
        self.location_c2 = distribution(self.c2_targets) if c2_targets else None
        self.stats.update({"beacons": 0, "routed": 0, "c2_received": 0, "first_beacon_us": 0, "exfil_bytes": 0})
        gsrandom.set_default_rng(self.rng)
        try:
            self.www = Network(self, "0.0.0.0/0", latency=bounded(expo(1 * S), lower=20 * MS),
                               bandwidth=expo(10 * MbPS))
            if infected_index is None:
                infected = next(distribution(self.workstations))        # malware_simple.py:114
            else:
                infected = self.workstations[infected_index]            # backdoor.py:69
        finally:
            gsrandom.set_default_rng(None)
        self.infected = infected
        infected.install(PROG_CALL_HOME, self.call_home)
        infected.install(PROG_NO_GOOD, self.no_good)
        self.router_local_addr = self.router.address_default
        self.router_inet_addr = self.router.link_to(self.www)
        self.router.install(PROG_ROUTE, self.route)
        self.bears = []
        for k, (addr, port) in enumerate(self.c2_targets):
            bear = Node(self, "Mama Bear" if k == 0 else f"Mama Bear {k + 1}", self.www, addr)
            bear.install(PROG_C2, self.command_and_control, port)
            self.bears.append(bear)
        self.bear = self.bears[0]

    @tagged(Tag.MALWARE)
    def call_home(self, ws: Node) -> None:
        """malware_simple.py:53-58 (``@malware`` = ``tagged(Tag.MALWARE)``, itsim/__init__.py:52-57)."""
        while True:
            advance(next(self.beacon_gap if self.beacon_gap is not None else expo(10)))
            proc = add(self.exfiltrate, ws)
            tr.Tracer.label(proc, tr.ROLE_MODEL, PROG_EXFILTRATE, ws.index)

    @tagged(Tag.MALWARE)
    def no_good(self, ws: Node) -> None:
        """malware_simple.py:61-65."""
        while True:
            advance(next(self.privileged_gap if self.privileged_gap is not None else expo(100)))
            proc = add(self.sudo_transmit, ws)
            tr.Tracer.label(proc, tr.ROLE_MODEL, PROG_SUDO_TRANSMIT, ws.index)

    @tagged(Tag.VULNERABLE)
    def sudo_transmit(self, ws: Node) -> None:
        """malware_simple.py:68-72."""
        proc = add(self.exfiltrate, ws)
        tr.Tracer.label(proc, tr.ROLE_MODEL, PROG_EXFILTRATE, ws.index)

    def exfiltrate(self, ws: Node) -> None:
        """malware_simple.py:75-82."""
        self.stats["beacons"] += 1
        if self.stats["first_beacon_us"] == 0:
            self.stats["first_beacon_us"] = int(now() * 1.0e6)
        target = next(self.location_c2) if self.location_c2 is not None else self.c2_targets[0]   # backdoor.py:36
        size = next(self.beacon_size) if self.beacon_size is not None else 100                      # backdoor.py:37
        with ws.open_socket(self.BAD_NEWS_BEAR_PORT) as socket:
            socket.send((self.router_local_addr, self.LOCAL_ROUTER_PORT), size,
                        {"content": "exfil", "dest_addr": target})

    def route(self, router: Node) -> None:
        """malware_simple.py:96-107: the one-way router."""
        with router.open_socket((self.router_local_addr, self.LOCAL_ROUTER_PORT)) as i_socket:
            with router.open_socket((self.router_inet_addr, self.LOCAL_ROUTER_PORT)) as world_socket:
                while True:
                    msg = i_socket.recv()
                    self.stats["routed"] += 1
                    self.stats["exfil_bytes"] += msg.byte_size
                    world_socket.send(msg.payload["dest_addr"], msg.byte_size, msg.payload)

    @tagged(Tag.MALWARE)
    def command_and_control(self, bear: Node, port: int = BAD_NEWS_BEAR_PORT) -> None:
        """malware_simple.py:85-92."""
        with bear.open_socket(port) as socket:
            while True:
                socket.recv()
                self.stats["c2_received"] += 1


def run_replica(seed: int, replica: int, duration: float, trace: bool = True, trace_limit: Optional[int] = None,
                num_endpoints: Optional[int] = None, conn_window: Optional[int] = None) -> Model:
    """Builds and runs one replica for ``duration`` simulated seconds; caller reads ``.tracer`` / ``.stats``."""
    model = Model(seed, replica, num_endpoints=num_endpoints, trace=trace, trace_limit=trace_limit,
                  conn_window=conn_window)
    model.run(duration)
    return model


if __name__ == "__main__":
    import json
    import time
    dur = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    t0 = time.time()
    m = run_replica(0, 0, dur, trace=False)
    wall = time.time() - t0
    out = m.tracer.summary()
    out.update(m.stats)
    out.update({"wall_s": wall, "events_per_s": out["events"] / wall, "draws": m.rng.draws,
                "connections": len(m.connections)})
    print(json.dumps(out))
    m.close()
