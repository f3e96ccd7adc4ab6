"""
Shipped-``itsim`` modelling objects, compiled for the engine: ``Link``, ``Node`` / ``Endpoint`` / ``Router``,
``Interface``, routes, and the ``run_proc`` family.

Reference signatures are kept (``itsim/network/link.py:37-42``, ``itsim/machine/node.py:84-133,279-303``,
``itsim/network/route.py:8-59``, ``itsim/types.py:32-72``); the one difference is the same as everywhere in this
package: a process body is the *name of a program* assembled with :class:`itsim_b200.ir.Asm`
(``Asm.thread_program``), not a Python function.
"""
import sys
from ipaddress import IPv4Address, ip_address, ip_network
from typing import Any, List, Optional, Sequence, Union

import numpy as np

from . import _lib as L
from .model import Compiled, Simulator, NET_LINK
from .random import Dist, bounded, constant
from .units import GbPS

AddressRepr = Union[None, str, int, IPv4Address]

ROUTE_LOCAL, ROUTE_RELAY = 0, 1
INIT_ADD, INIT_RUN_PROC = 0, 1

_creation_counter = [0]


def as_address(ar: AddressRepr, rr: Any = "0.0.0.0/0") -> IPv4Address:
    """``itsim/types.py:32-72``: the host bits of ``ar`` re-rooted into network ``rr``."""
    network = rr if hasattr(rr, "network_address") else ip_network(rr)
    if ar is None:
        machine = ip_address(0)
    elif isinstance(ar, int):
        if ar < 0 or ar >= 2 ** 32:
            raise ValueError(ar)
        machine = ip_address(ar)
    elif isinstance(ar, str):
        machine = ip_address(ar)
    else:
        machine = ar
    return network.network_address + (int(machine) & int(network.hostmask))


class Route:
    """``itsim/network/route.py:8-32``."""
    kind = ROUTE_LOCAL

    def __init__(self, cr: Optional[str] = None) -> None:
        self.cidr = ip_network(cr or "0.0.0.0/0")
        self.gateway = ip_address(0)


class Local(Route):
    """Deliver directly (``route.py:35-41``)."""


class Relay(Route):
    """Through a gateway (``route.py:44-59``)."""
    kind = ROUTE_RELAY

    def __init__(self, gr: AddressRepr, cr: Optional[str] = None) -> None:
        super().__init__(cr)
        self.gateway = as_address(gr)


class Link:
    """``Link(c, latency, bandwidth)`` (``itsim/network/link.py:25-42``): latency clipped at 0, bandwidth at the
    smallest positive double; members are kept in connection order (the canonicalisation of the reference's
    hash-ordered set)."""

    def __init__(self, c: str, latency: Dist, bandwidth: Dist) -> None:
        self.cidr = ip_network(c) if isinstance(c, str) else c
        self._latency = bounded(latency, lower=0.0)
        self._bandwidth = bounded(bandwidth, lower=sys.float_info.min)
        self._nodes: List["Node"] = []

    def iter_nodes(self):
        return iter(self._nodes)

    def _connect(self, node: "Node") -> None:
        if node not in self._nodes:
            self._nodes.append(node)


class Loopback(Link):
    """``itsim/network/link.py:89-92``."""

    def __init__(self) -> None:
        super().__init__("127.0.0.0/8", constant(0), constant(100 * GbPS))


class Interface:
    """``itsim/network/interface.py:22-68``: routes always start with ``Local(link.cidr)``."""

    def __init__(self, link: Link, address: IPv4Address, routes: List[Route]) -> None:
        self.link = link
        self._routes = routes
        self.address = as_address(address, link.cidr)

    @property
    def cidr(self):
        return self.link.cidr

    @property
    def routes(self) -> List[Route]:
        first = Local()
        first.cidr = self.link.cidr
        return [first] + self._routes


class Node:
    """``itsim/machine/node.py:69-133``: interfaces by link, loopback first."""

    def __init__(self) -> None:
        self._order = _creation_counter[0]
        _creation_counter[0] += 1
        self._interfaces: List[Interface] = []
        self.index = -1
        self._connect_to(Loopback(), "127.0.0.1")

    def _connect_to(self, link: Link, ar: AddressRepr = None, routes: Optional[List[Route]] = None) -> Interface:
        link._connect(self)
        interface = Interface(link, as_address(ar, link.cidr), routes or [])
        for i, existing in enumerate(self._interfaces):      # _interfaces is keyed by cidr in the reference
            if existing.link.cidr == link.cidr:
                self._interfaces[i] = interface
                return interface
        self._interfaces.append(interface)
        return interface

    def connected_to_static(self, link: Link, ar: AddressRepr, routes: Optional[List[Route]] = None) -> "Node":
        self._connect_to(link, ar, routes)
        return self

    def addresses(self):
        return (interface.address for interface in self._interfaces)

    def interfaces(self):
        return iter(self._interfaces)

    def run_proc_in(self, sim: Simulator, time: float, program: str, *regs: int) -> "Node":
        """``node.py:279-283``: new Process (pid from the node's counter), first Thread, body started through
        ``Thread.run_in(time, ...)``."""
        sim._machine_procs.append((self, float(time), program, regs))
        return self

    def run_proc(self, sim: Simulator, program: str, *regs: int) -> "Node":
        return self.run_proc_in(sim, 0, program, *regs)

    with_proc_in = run_proc_in
    with_proc = run_proc

    def schedule_daemon_in(self, sim: Simulator, time: float, program: str, *regs: int) -> "Node":
        """``node.py:352-357``: a daemon's trigger run as a new process after ``time``."""
        return self.run_proc_in(sim, time, program, *regs)

    def connected_to_dynamic(self, link: Link, sim: Simulator, program: str = "dhcp_client") -> "Node":
        """``node.py:135-154``: connect at host number 0 of the link's CIDR and start a DHCP client on the new
        interface (``program`` = the assembled ``DHCPClient`` body, ``machine_models.dhcp_exchange``)."""
        self._connect_to(link, as_address(None), None)
        return self.schedule_daemon_in(sim, 0.0, program)


class Endpoint(Node):
    """``itsim/machine/endpoint.py:4-8``."""


class Router(Node):
    """``itsim/network/router.py:4-15``.  The reference's Router is an empty Node (packets in transit are dropped,
    node.py:245-250); ``forwarding=True`` installs the rule the reference documents as intent for
    ``handle_packet_transit`` (sphinx/source/networks.rst:275-280): ``interface, hop = _solve_transfer(dest);
    interface.link._transfer_packet(packet, hop)``."""

    def __init__(self, forwarding: bool = False) -> None:
        super().__init__()
        self.forwarding = forwarding


def compile_machine(sim: Simulator) -> Compiled:
    """Tables for a model built from the classes above (every node a process was started on, plus every node sharing
    a link with one of them)."""
    sim.asm.add_hidden_programs()
    code, entries_raw = sim.asm.assemble()
    entries = np.zeros(len(entries_raw), dtype=L.ENTRY_DTYPE)
    entries["pc"], entries["prog"] = entries_raw[:, 0], entries_raw[:, 1]

    # closure of nodes over links, numbered in creation order
    nodes: List[Node] = []
    pending = [n for n, _, _, _ in sim._machine_procs]
    while pending:
        node = pending.pop()
        if node in nodes:
            continue
        nodes.append(node)
        for interface in node.interfaces():
            pending.extend(interface.link.iter_nodes())
    nodes.sort(key=lambda n: n._order)
    for i, node in enumerate(nodes):
        node.index = i

    links: List[Link] = []
    for node in nodes:
        for interface in node.interfaces():
            if interface.link not in links:
                links.append(interface.link)

    members: List[int] = []
    nets = np.zeros(len(links), dtype=L.NET_DTYPE)
    for k, link in enumerate(links):
        first = len(members)
        members.extend(n.index for n in link.iter_nodes())
        nets[k] = (int(link.cidr.network_address), int(link.cidr.netmask), int(link.cidr.broadcast_address), first,
                   len(members) - first, sim.dist(link._latency), sim.dist(link._bandwidth), NET_LINK)

    ifaces, routes = [], []
    node_rows = np.zeros(len(nodes), dtype=L.NODE_DTYPE)
    for node in nodes:
        first_iface = len(ifaces)
        for interface in node.interfaces():
            first_route = len(routes)
            for route in interface.routes:
                routes.append((int(route.cidr.network_address), int(route.cidr.netmask), route.cidr.prefixlen,
                               route.kind, int(route.gateway), (0, 0, 0)))
            ifaces.append((node.index, links.index(interface.link), int(interface.address), first_route,
                           len(routes) - first_route, (0, 0, 0)))
        n_if = len(ifaces) - first_iface
        visible = [i for i in node.interfaces() if not i.address.is_loopback]
        node_rows[node.index] = (int((visible or list(node.interfaces()))[0].address),
                                 links.index(list(node.interfaces())[0].link),
                                 1 if getattr(node, "forwarding", False) else 0, first_iface | (n_if << 16))

    init = np.zeros(len(sim._machine_procs), dtype=L.INIT_DTYPE)
    for i, (node, delay, program, regs) in enumerate(sim._machine_procs):
        r = list(regs) + [0] * (4 - len(regs))
        init[i] = (sim.asm.entry(program), node.index, r, INIT_RUN_PROC, sim.const(delay))

    dists = np.zeros(len(sim._dists), dtype=L.DIST_DTYPE)
    for i, d in enumerate(sim._dists):
        dists[i] = (d.kind, d.flags, d.p0, d.p1, d.slope, d.shift, d.lower, d.upper)
    consts = np.array(sim._consts, dtype=np.float64)
    filters = np.zeros(0, dtype=L.FILTER_DTYPE)
    caps = dict(sim.caps)
    for key, default in (("max_threads", 64), ("max_tasks", 64), ("max_signals", 256)):
        if not caps.get(key):
            caps[key] = default
    return Compiled(code, entries, dists, consts, nets, node_rows, init, filters, L.Caps(**caps),
                    ifaces=np.array(ifaces, dtype=L.IFACE_DTYPE), routes=np.array(routes, dtype=L.ROUTE_DTYPE),
                    members=np.array(members, dtype=np.uint32))
