"""
Host-side model compiler: ITsim's modelling objects, compiled to flat struct-of-arrays tables for the engine.

The classes keep the reference's names, constructor arguments and error behaviour for the part of the API the hot
path uses, so a model script reads like the reference's:

==========================  =====================================================================================
here                        reference
==========================  =====================================================================================
``Simulator``               ``itsim.simulator.Simulator`` (``itsim/simulator.py:9``): ``run(duration)``, ``now()``
``Network``                 override-path ``Network`` (``examples/network_simple/network_simple_overrides.py:189-297``)
``Endpoint`` / ``install``  ``Endpoint`` (``overrides.py:403-433``)
``Workstation``             ``Workstation`` (``examples/network_simple/network_simple.py:42-98``)
``Dist`` factories          ``greensim.random`` / ``itsim.random.num_bytes`` (``itsim_b200/random.py``)
==========================  =====================================================================================

What cannot be kept is the process *body*: the reference runs arbitrary Python functions as coroutines
(SURVEY.md F6).  Here a body is a named program assembled once per model with :class:`itsim_b200.ir.Asm`;
``install`` takes the program name instead of a function.  New on top of the reference API: ``run_replicas``,
the batched entry that advances N independent replicas (Philox key = (seed, replica)) on the GPU.
"""
from ipaddress import ip_network
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .ir import Asm
from .random import Dist

NET_OVERRIDE, NET_LINK = 0, 1


class AddressInUse(Exception):
    """``itsim/network/location.py`` AddressInUse (raised by ``Network.link``, overrides.py:242-243)."""


class InvalidAddress(Exception):
    """``itsim/network/location.py`` InvalidAddress (overrides.py:240-241)."""


class NetworkFull(Exception):
    """overrides.py:436-437."""


class Compiled:
    """The tables handed to ``itsb_load_model`` (blob order = ``ITSB_BLOB_*`` in include/itsb.h)."""

    def __init__(self, code, entries, dists, consts, nets, nodes, init, filters, caps: L.Caps, ifaces=None,
                 routes=None, members=None, address_tags: int = 0) -> None:
        self.address_tags = address_tags
        self.code, self.entries, self.dists, self.consts = code, entries, dists, consts
        self.nets, self.nodes, self.init, self.filters, self.caps = nets, nodes, init, filters, caps
        self.ifaces = ifaces if ifaces is not None else np.zeros(0, dtype=L.IFACE_DTYPE)
        self.routes = routes if routes is not None else np.zeros(0, dtype=L.ROUTE_DTYPE)
        self.members = members if members is not None else np.zeros(0, dtype=np.uint32)

    def blobs(self) -> List[np.ndarray]:
        return [self.code, self.entries, self.dists, self.consts, self.nets, self.nodes, self.init, self.filters,
                self.ifaces, self.routes, self.members]

    def desc(self) -> L.ModelDesc:
        return L.ModelDesc(L.ABI_VERSION, len(self.code), len(self.entries), len(self.dists), len(self.consts),
                           len(self.nets), len(self.nodes), len(self.init), len(self.filters), len(self.ifaces),
                           len(self.routes), len(self.members), self.address_tags, self.caps)


class Simulator:
    """Build-time container and run-time handle.  ``run``/``now`` follow greensim's ``Simulator``; the model is
    compiled on first use and the replica state then lives on the device (repeated ``run`` continues)."""

    def __init__(self, ts_now: float = 0.0, name: Optional[str] = None) -> None:
        if ts_now != 0.0:
            raise ValueError("the engine starts replicas at t = 0")
        self.name = name
        self.asm = Asm()
        self._dists: List[Dist] = []
        self._dist_index: Dict[tuple, int] = {}
        self._consts: List[float] = []
        self._filters: List[Tuple[int, int, int, int]] = []
        self.networks: List["Network"] = []
        self.nodes: List["Endpoint"] = []
        self._installed: List[Tuple[str, "Endpoint", Sequence[int]]] = []
        self.caps = dict(max_events=384, max_procs=256, max_sockets=128, max_packets=0, arena_nodes=16384,
                         trace_records=0, max_threads=0, max_tasks=0, max_signals=0, table_words=0, net_event_records=0,
                         connection_records=0, connection_window=8)
        self.address_tags = 0       # payload tags whose f1 field is an ADDRESS entry (connection records)
        self._machine_procs: List[Tuple[Any, float, str, Sequence[int]]] = []   # Node.run_proc_in calls, in order
        self._engine: Any = None
        self._lib: Any = None       # tests inject the host-emulation library here; None = the CUDA engine

    # -- tables ------------------------------------------------------------------------------------------------
    def dist(self, d: Dist) -> int:
        key = d.key()
        if key not in self._dist_index:
            self._dist_index[key] = len(self._dists)
            self._dists.append(d)
        return self._dist_index[key]

    def const(self, value: float) -> int:
        value = float(value)
        if value not in self._consts:
            self._consts.append(value)
        return self._consts.index(value)

    def packet_filter(self, always: Sequence[int] = (), if_host: Sequence[int] = (),
                      if_list: Sequence[int] = (), if_listed: Sequence[int] = ()) -> int:
        """Declares which payload tags a ``recv_filter`` loop hands to its program: ``always``; ``if_host`` only when
        the packet's hostname field is the node's own; ``if_list`` only while the process remembers a packet;
        ``if_listed`` only while it remembers one with the packet's hostname field (the entry ``list_take`` on that
        field would find).  Every other packet is consumed by the loop head, exactly as the body's own dispatch would
        have ignored it."""
        def mask(tags: Sequence[int]) -> int:
            m = 0
            for t in tags:
                if not 0 <= t < 32:
                    raise ValueError("payload tags used in filters must be below 32")
                m |= 1 << t
            return m
        entry = (mask(always), mask(if_host), mask(if_list), mask(if_listed))
        if entry not in self._filters:
            self._filters.append(entry)
        return self._filters.index(entry)

    def compile(self) -> Compiled:
        if self._machine_procs:
            from .machine import compile_machine
            return compile_machine(self)
        consts = np.array(self._consts, dtype=np.float64)
        # nodes are numbered network by network, in link order, so a network's members are one index range
        order: List[Endpoint] = []
        nets = np.zeros(len(self.networks), dtype=L.NET_DTYPE)
        for k, net in enumerate(self.networks):
            first = len(order)
            order.extend(net._members)
            nets[k] = (int(net.cidr.network_address), int(net.cidr.netmask), int(net.cidr.broadcast_address), first,
                       len(net._members), self.dist(net._latency), self.dist(net._bandwidth), net.mode)
        if len(order) != len(self.nodes):
            raise ValueError("every node must be linked to exactly one network")
        for idx, node in enumerate(order):
            node.index = idx
        code, entries_raw = self.asm.assemble()     # after numbering: programs may name nodes
        entries = np.zeros(len(entries_raw), dtype=L.ENTRY_DTYPE)
        entries["pc"], entries["prog"] = entries_raw[:, 0], entries_raw[:, 1]
        nodes = np.zeros(len(order), dtype=L.NODE_DTYPE)
        for node in order:
            nodes[node.index] = (node.address, self.networks.index(node.network), node.host_id, 0)
        init = np.zeros(len(self._installed), dtype=L.INIT_DTYPE)
        for i, (program, node, regs) in enumerate(self._installed):
            r = list(regs) + [0] * (4 - len(regs))
            if isinstance(node, DrawnNode):
                # the engine adds the drawn index to the first candidate's node index (replica_setup, itsb_core.cuh)
                if [c.index for c in node.candidates] != list(range(node.first.index, node.first.index + len(node.candidates))):
                    raise ValueError("the candidates of a drawn node must be numbered consecutively (link them to "
                                     "their network one after the other)")
                init[i] = (self.asm.entry(program), node.first.index, r, node.mode, self.dist(node.dist))
            else:
                init[i] = (self.asm.entry(program), node.index, r, 0, 0)
        filters = np.zeros(len(self._filters), dtype=L.FILTER_DTYPE)
        for i, (fa, fh, fl, fm) in enumerate(self._filters):
            filters[i] = (fa, fh, fl, fm)
        dists = np.zeros(len(self._dists), dtype=L.DIST_DTYPE)   # last: networks and drawn installs add generators
        for i, d in enumerate(self._dists):
            dists[i] = (d.kind, d.flags, d.p0, d.p1, d.slope, d.shift, d.lower, d.upper)
        return Compiled(code, entries, dists, consts, nets, nodes, init, filters, L.Caps(**self.caps),
                        address_tags=self.address_tags)

    # -- running -------------------------------------------------------------------------------------------------
    def run_replicas(self, n: int, seed: int = 0, duration: Optional[float] = None, first_replica_id: int = 0,
                     device: int = 0) -> Any:
        """Compiles the model, creates ``n`` replicas on ``device`` and (optionally) runs them for ``duration``."""
        from .engine import Engine
        self._engine = Engine(self.compile(), device=device, lib=self._lib)
        self._engine.init_replicas(n, seed, first_replica_id)
        if duration is not None:
            self._engine.run(duration)
        return self._engine

    def run(self, duration: float, seed: int = 0) -> None:
        """``Simulator.run(duration)`` for a single replica (replica id 0 of ``seed``)."""
        if self._engine is None:
            self.run_replicas(1, seed)
        self._engine.run(duration)

    def now(self) -> float:
        return float(self._engine.now(0, 1)[0]) if self._engine is not None else 0.0


class DrawnNode:
    """A node chosen per replica at set-up with one draw from the replica's stream:
    ``infected = next(distribution(workstation_list))`` (malware_simple.py:114).  ``install`` on it installs on
    whichever node the replica drew; a second ``install`` lands on the same node."""

    def __init__(self, sim: "Simulator", candidates: Sequence["Endpoint"]) -> None:
        from .random import distribution_index
        self.sim = sim
        self.first = candidates[0]
        self.candidates = list(candidates)
        self.dist = distribution_index(len(candidates))
        self.mode = 2          # ITSB_INIT_ADD_DRAWN for the first install, ITSB_INIT_ADD_SAME afterwards
        self._candidates = list(candidates)

    def install(self, program: str, *regs: int) -> None:
        drawn = DrawnNode.__new__(DrawnNode)
        drawn.__dict__.update(self.__dict__)
        self.sim._installed.append((program, drawn, regs))
        self.mode = 3


class Network:
    """One flat broadcast domain (overrides.py:189-297).  ``latency`` / ``bandwidth`` are clipped like the reference
    clips them: ``bounded(latency, lower=0)``, ``bounded(bandwidth, lower=0.125)`` (overrides.py:202-203)."""

    mode = NET_OVERRIDE

    def __init__(self, sim: Simulator, cidr: str, latency: Dist, bandwidth: Dist) -> None:
        from .random import bounded
        self.sim = sim
        self.cidr = ip_network(cidr)
        self._latency = bounded(latency, lower=0)
        self._bandwidth = bounded(bandwidth, lower=0.125)
        self._members: List["Endpoint"] = []
        self._addresses: Dict[int, "Endpoint"] = {}
        self._top_free = (int(a) for a in self.cidr.hosts())
        sim.networks.append(self)

    @property
    def address_broadcast(self) -> int:
        return int(self.cidr.broadcast_address)

    def link(self, node: "Endpoint", ar: Optional[Any] = None) -> int:
        if ar is None:
            try:
                address = next(self._top_free)
            except StopIteration:
                raise NetworkFull()
        else:
            address = int(ar) if isinstance(ar, int) else int(ip_network(f"{ar}/32").network_address)
        if not (int(self.cidr.network_address) <= address <= self.address_broadcast):
            raise InvalidAddress(address)
        if address in self._addresses:
            raise AddressInUse(address)
        self._addresses[address] = node
        self._members.append(node)
        return address


class Endpoint:
    """overrides.py:403-433.  ``install(program, ...)`` = ``sim.add(fn_software, self, ...)`` at t = 0."""

    def __init__(self, name: str, network: Network, ar: Optional[Any] = None, host_id: int = 0) -> None:
        self.name = name
        self.network = network
        self.sim = network.sim
        self.host_id = host_id
        self.index = -1
        self.address = network.link(self, ar)
        self.sim.nodes.append(self)

    @property
    def address_default(self) -> int:
        return self.address

    def install(self, program: str, *regs: int) -> None:
        self.sim._installed.append((program, self, regs))

    def link_to(self, network: Network, ar: Optional[Any] = None) -> "Endpoint":
        """``Node.link_to`` (overrides.py:307-312): a further address on another network.  The attachment is
        numbered like a node of its own (a machine's attachments share nothing the engine needs to know about)."""
        return Endpoint(f"{self.name}@{network.cidr}", network, ar, host_id=self.host_id)


class Workstation(Endpoint):
    """network_simple.py:42-98: an endpoint with a wake signal; the signal itself lives in the replica state."""
