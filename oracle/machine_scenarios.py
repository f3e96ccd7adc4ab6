"""
ORACLE / TEST INFRASTRUCTURE ONLY -- golden traces for the shipped-``itsim`` primitives.

Runs the reference's **own integration-test bodies, unmodified** (imported from /root/reference/tests/integ/*.py) on
the reference's unmodified ``itsim`` package on top of the greensim shim, under the event tracer, with the
canonicalisations of SURVEY.md section 8-c applied by monkey-patching at run time:

* ``Link._nodes``, ``Process._threads`` / ``_children``, ``Thread._computations`` become insertion-ordered sets
  (the reference uses hash-ordered ``set``s of id-hashed objects: F9);
* every generator draws from one Philox stream (seed, replica) through the shim's default-rng hook.

Only usable where /root/reference exists (the build container).  ``oracle/make_golden.py --machine`` stores the
traces under tests/golden/machine_*.npz; the engine-side restatements of the same bodies are in
``itsim_b200/machine_models.py``.
"""
import collections.abc
import importlib.util
import os
import sys
from typing import Any, Callable, Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ITSIM_REFERENCE", "/root/reference")
for _p in (REFERENCE, _HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import greensim  # noqa: E402  (the shim)
import greensim.random as gsrandom  # noqa: E402
import evtrace as tr  # noqa: E402
from philox import PhiloxRandom  # noqa: E402

MARK = 9                     # trace-only record: a model-visible side effect (cemetary.append(name), ledger.add(name))
PROG_DELIVER = 10
EXC_CODES = {"ThreadKilled": 0x81, "Timeout": 0x82}


class OrderedSet(collections.abc.MutableSet):
    """Insertion-ordered stand-in for the reference's ``set`` containers."""

    def __init__(self, items=()):
        self._d = dict.fromkeys(items)

    def __contains__(self, x): return x in self._d
    def __iter__(self): return iter(list(self._d))
    def __len__(self): return len(self._d)
    def add(self, x): self._d[x] = None
    def discard(self, x): self._d.pop(x, None)

    def remove(self, x):
        del self._d[x]

    def __ior__(self, other):
        for x in other:
            self.add(x)
        return self

    def __isub__(self, other):
        for x in list(other):
            self.discard(x)
        return self

    def __eq__(self, other):
        return set(self._d) == set(other)

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return f"OrderedSet({list(self._d)})"


_patched = False


def canonicalise_reference() -> None:
    global _patched
    if _patched:
        return
    _patched = True
    from itsim.network import link as link_mod
    from itsim.machine.process_management import process as process_mod, thread as thread_mod

    def wrap(cls, attrs):
        original = cls.__init__

        def init(self, *a, **k):
            original(self, *a, **k)
            for attr in attrs:
                setattr(self, attr, OrderedSet(getattr(self, attr)))
        cls.__init__ = init

    wrap(link_mod.Link, ["_nodes"])
    wrap(process_mod.Process, ["_threads", "_children"])
    wrap(thread_mod.Thread, ["_computations"])


def load_reference_test(rel_path: str):
    path = os.path.join(REFERENCE, rel_path)
    spec = importlib.util.spec_from_file_location("ref_" + os.path.basename(rel_path)[:-3], path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


class Scenario:
    """One traced run.  ``programs`` maps body function names to program ids; nodes are numbered in creation order."""

    def __init__(self, programs: Dict[str, int], seed: int = 0, replica: int = 0) -> None:
        canonicalise_reference()
        from itsim.simulator import Simulator
        self.sim = Simulator()
        self.rng = PhiloxRandom(seed, replica)
        self.programs = programs
        self.nodes: List[Any] = []
        self.tracer = tr.Tracer(self.sim)
        self.tracer.exc_code = lambda exc: EXC_CODES.get(type(exc).__name__, 0)
        self.tracer.deliver_aux = self._deliver_aux
        self.tracer.resolver = self._resolve

    def node(self, n: Any) -> Any:
        self.nodes.append(n)
        return n

    def _node_index(self, n: Any) -> int:
        for i, m in enumerate(self.nodes):
            if m is n:
                return i
        return 0xFFFF

    def _resolve(self, proc: Any) -> tr.Ident:
        body = proc._body
        owner = getattr(body, "__self__", None)
        if owner is not None and body.__name__ == "_receive_packet":
            return tr.Ident(tr.ROLE_DELIVER, PROG_DELIVER, self._node_index(owner))
        cells = {}
        if getattr(body, "__closure__", None):
            cells = dict(zip(body.__code__.co_freevars, (c.cell_contents for c in body.__closure__)))
        if body.__name__ == "wrap_computation" and "self" in cells:
            thread = cells["self"]
            f = cells.get("f")
            name = getattr(f, "__name__", "?")
            return tr.Ident(tr.ROLE_MODEL, self.programs.get(name, 0), self._node_index(thread._process.node))
        return tr.Ident(tr.ROLE_MODEL, self.programs.get(body.__name__, 0), 0xFFFF)

    def _deliver_aux(self, proc: Any, ev: Any) -> int:
        node, packet = proc._body.__self__, ev.args[0]
        dest = packet.dest.hostname_as_address()
        mine = any(i.address == dest or i.cidr.broadcast_address == dest for i in node.interfaces())
        accepted = 1 if (mine and packet.dest.port in node._sockets) else 0
        return (int(packet.byte_size) & 0x7FFFFFFF) | (accepted << 31)

    def mark(self, code: int, node: int = 0) -> None:
        self.tracer.records.append((self.sim.now(), MARK, 0, node, code))

    def run(self, duration=None) -> None:
        gsrandom.set_default_rng(self.rng)
        try:
            if duration is None:
                self.sim.run()
            else:
                self.sim.run(duration)
        finally:
            gsrandom.set_default_rng(None)

    def result(self):
        arr = self.tracer.as_array()
        out = self.tracer.summary()
        out.update({"now": self.sim.now(), "draws": self.rng.draws, "digest": tr.trace_digest(arr)})
        self.sim._clear()
        return arr, out


class MarkList(list):
    def __init__(self, scenario: Scenario, codes: Dict[str, int]) -> None:
        super().__init__()
        self._scenario, self._codes = scenario, codes

    def append(self, name):
        super().append(name)
        self._scenario.mark(self._codes[name])


class MarkSet(set):
    def __init__(self, scenario: Scenario, codes: Dict[str, int]) -> None:
        super().__init__()
        self._scenario, self._codes = scenario, codes

    def add(self, name):
        super().add(name)
        self._scenario.mark(self._codes[name])


# --------------------------------------------------------------------------------------------------------------------
# scenarios = the reference's integration tests
# --------------------------------------------------------------------------------------------------------------------
PROGS_PACKET_TRANSFER = {"client": 1, "server": 2}
MARKS_PACKET_TRANSFER = {"client": 1, "server": 2}


def packet_transfer(seed: int = 0, replica: int = 0):
    """tests/integ/test_packet_transfer.py:16-55."""
    T = load_reference_test("tests/integ/test_packet_transfer.py")
    from greensim.random import uniform, constant
    from itsim.machine.endpoint import Endpoint
    from itsim.network.link import Link
    from itsim.units import S, MS, MbPS
    sc = Scenario(PROGS_PACKET_TRANSFER, seed, replica)
    T.ledger = MarkSet(sc, MARKS_PACKET_TRANSFER)
    link = Link("10.11.12.0/24", latency=uniform(100 * MS, 200 * MS), bandwidth=constant(100 * MbPS))
    pinger = sc.node(Endpoint().connected_to_static(link, "10.11.12.10"))
    pinger.run_proc_in(sc.sim, 0.1, T.client)
    ponger = sc.node(Endpoint().connected_to_static(link, "10.11.12.20"))
    ponger.run_proc(sc.sim, T.server)
    sc.run(10.0 * S)
    assert T.ledger == {"client", "server"}
    return sc.result()


PROGS_BROADCAST = {"server": 1, "client": 2}


def broadcast(seed: int = 0, replica: int = 0, duration: float = 10.0):
    """tests/integ/test_broadcast.py:17-53."""
    T = load_reference_test("tests/integ/test_broadcast.py")
    from greensim.random import uniform, constant
    from itsim.network.link import Link
    from itsim.types import as_address
    from itsim.units import MS, MbPS
    sc = Scenario(PROGS_BROADCAST, seed, replica)
    link = Link("10.11.0.0/16", uniform(100 * MS, 200 * MS), constant(100 * MbPS))
    endpoints = []
    for n in T.NUMS_HOST:
        # EndpointChattering.__init__ starts its two processes before the node is connected to the link
        ep = T.EndpointChattering(sc.sim)
        sc.node(ep)
        endpoints.append(ep.connected_to_static(link, as_address(n, link.cidr)))
    sc.run(duration)
    all_addresses = set(as_address(n, link.cidr) for n in T.NUMS_HOST)
    if duration >= 10.0:
        for endpoint in endpoints:
            endpoint.assert_seen_all_peers(all_addresses)
    return sc.result()


PROGS_THREADS = {"parent": 1, "watcher": 2, "elder": 3, "cadet": 4, "super_long": 5, "grandchild": 6}
MARKS_THREADS = {"cadet": 1, "grandchild": 2, "elder": 3, "parent": 4, "super_long": 5, "watcher": 6}


def thread_lifecycle():
    """tests/integ/test_thread_lifecycle.py:21-84 -- exact order and final clock are the reference's known answer."""
    T = load_reference_test("tests/integ/test_thread_lifecycle.py")
    from itsim.machine.endpoint import Endpoint
    sc = Scenario(PROGS_THREADS)
    cemetary = MarkList(sc, MARKS_THREADS)
    sc.node(Endpoint()).with_proc_in(sc.sim, 0, T.parent, cemetary)
    sc.run()
    assert list(cemetary) == ["cadet", "grandchild", "elder", "parent", "super_long", "watcher"]
    assert sc.sim.now() == 120
    return sc.result()


PROGS_PROCESSES = {"adameve": 1, "life": 2}
MARKS_PROCESSES = {"Cain": 1, "Abel": 2, "Seth": 3, "Humanity": 4, "adameve": 5}


def process_lifecycle():
    """tests/integ/test_process_lifecycle.py:41-112."""
    T = load_reference_test("tests/integ/test_process_lifecycle.py")
    from itsim.machine.endpoint import Endpoint
    sc = Scenario(PROGS_PROCESSES)
    cemetary = MarkSet(sc, MARKS_PROCESSES)
    sc.node(Endpoint()).with_proc(sc.sim, T.adameve, cemetary)
    sc.run(4000)
    assert set(cemetary) == {"adameve", "Cain", "Abel", "Seth"}
    return sc.result()


PROGS_DHCP = {"forward_recv": 1, "trigger": 2}


def dhcp_exchange(seed: int = 0, replica: int = 0, duration: float = 10.0, clients: int = 3):
    """tests/integ/test_dhcp_exchange.py:17-38 -- the reference's DHCPServer behind a networking daemon on a Router,
    `clients` endpoints running the reference's DHCPClient; known answer: they end up with hosts 100, 101, 102."""
    from greensim.random import uniform, constant
    from itsim.machine.endpoint import Endpoint
    from itsim.network.link import Link
    from itsim.network.router import Router
    from itsim.network.service.dhcp.client import DHCPClient
    from itsim.network.service.dhcp.server import DHCPServer
    from itsim.types import Protocol, as_address
    from itsim.units import MS, MbPS
    sc = Scenario(PROGS_DHCP, seed, replica)
    link = Link("10.1.128.0/18", uniform(100 * MS, 200 * MS), constant(100 * MbPS))
    router = sc.node(Router().connected_to_static(link, "10.1.128.1"))
    router.networking_daemon(sc.sim, Protocol.UDP, 67)(DHCPServer(100, link.cidr, as_address("10.1.128.1")))
    endpoints = [sc.node(Endpoint().connected_to_static(link, as_address(None))) for _ in range(clients)]
    for endpoint in endpoints:
        endpoint.schedule_daemon_in(sc.sim, 0.0, DHCPClient(endpoint._interfaces[link.cidr]))
    sc.run(duration)
    got = sorted(int(a) - int(link.cidr.network_address) for e in endpoints for a in e.addresses() if not a.is_loopback)
    arr, out = sc.result()
    out["hosts"] = got
    if clients == 3 and duration >= 10.0:
        assert got == [100, 101, 102], got
    return arr, out


PROGS_LAN = {"server": 1, "client": 2}
LAN_PORT, TAG_CHATTER, TAG_REQUEST, TAG_REPLY = 10000, 3, 4, 5


def large_lan(seed: int = 0, replica: int = 0, duration: float = 8.0, subnets: int = 16, per_subnet: int = 64):
    """Config C4: `subnets` access links (/25) of `per_subnet` endpoints, one router per access
    link, all routers on a backbone /24; endpoints reach other subnets through Relay default routes.  Built from the
    reference's own Link / Endpoint / Router / Relay classes; the router's forwarding is the two-line rule the
    reference documents as intent for ``Node.handle_packet_transit`` (sphinx/source/networks.rst:275-280, SURVEY
    section 8-f row 1).  Traffic: every endpoint runs a server (recv loop on port 10000 answering requests) and a client
    (every U(3, 6) s: a 128-byte broadcast on its link, then a 64-byte request to the same host number on the next
    subnet) -- the bodies of tests/integ/test_broadcast.py:28-40 and test_packet_transfer.py:16-47 combined."""
    from greensim.random import uniform, constant, normal
    from itsim.machine.endpoint import Endpoint
    from itsim.network.link import Link
    from itsim.network.route import Relay
    from itsim.network.router import Router
    from itsim.simulator import advance
    from itsim.types import Protocol, as_address
    from itsim.units import MS, GbPS

    class ForwardingRouter(Router):
        def handle_packet_transit(self, packet):
            interface, hop = self._solve_transfer(packet.dest.hostname_as_address())
            interface.link._transfer_packet(packet, hop)

    sc = Scenario(PROGS_LAN, seed, replica)
    interval = uniform(3.0, 6.0)

    def server(context):
        with context.node.bind(Protocol.UDP, LAN_PORT) as socket:
            while True:
                packet = socket.recv()
                if packet.payload.get("content") == "request":
                    socket.send(packet.source, 32, {"content": "reply"})

    def client(context, peer):
        with context.node.bind(Protocol.UDP) as socket:
            while True:
                advance(next(interval))
                for interface in context.node.interfaces():
                    if not interface.address.is_loopback:
                        socket.send((interface.cidr.broadcast_address, LAN_PORT), 128, {"content": "chatter"})
                socket.send((peer, LAN_PORT), 64, {"content": "request"})

    backbone = Link("10.30.0.0/24", normal(5 * MS, 1 * MS), constant(1 * GbPS))
    access = [Link(f"10.20.{i}.0/25", normal(5 * MS, 1 * MS), constant(1 * GbPS)) for i in range(subnets)]
    for i in range(subnets):
        router = ForwardingRouter()
        sc.node(router)
        router.connected_to_static(access[i], 1)
        router.connected_to_static(backbone, i + 1, [Relay(as_address(j + 1, backbone.cidr), f"10.20.{j}.0/25")
                                                    for j in range(subnets) if j != i])
    for i in range(subnets):
        for h in range(per_subnet):
            ep = Endpoint()
            sc.node(ep)
            ep.connected_to_static(access[i], h + 2, [Relay(as_address(1, access[i].cidr))])
            peer = as_address(h + 2, access[(i + 1) % subnets].cidr)
            ep.run_proc(sc.sim, server)
            ep.run_proc(sc.sim, client, peer)
    sc.run(duration)
    return sc.result()


# --------------------------------------------------------------------------------------------------------------------
# the front-end's test bodies (tests/frontend_bodies.py), run on the reference's itsim: the SAME source file that the
# engine's front-end compiles is executed here as ordinary Python coroutines on the reference's classes
# --------------------------------------------------------------------------------------------------------------------
FRONTEND_PROGS = {
    "ping_pong": {"pinger": 1, "ponger": 2},
    "chatter": {"listen": 1, "shout": 2},
    "threads": {"root": 1, "observer": 2, "first": 3, "second": 4, "sleeper": 5, "leaf": 6},
    "family": {"ancestors": 1, "life": 2},
    "lan": {"lan_server": 1, "lan_client": 2},
    "echo": {"answerer": 1, "asker": 2},
    "daemon": {"forward_recv": 1, "caller": 2},
}


def frontend_scenario(name: str, seed: int = 0, replica: int = 0, duration=10.0):
    """Returns (trace, summary); summary["state"] holds the host-side containers the bodies filled."""
    from itsim.network.router import Router
    path = os.path.join(os.path.dirname(_HERE), "tests", "frontend_bodies.py")
    spec = importlib.util.spec_from_file_location("frontend_bodies_on_reference", path)
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)

    class ForwardingRouter(Router):
        def handle_packet_transit(self, packet):
            interface, hop = self._solve_transfer(packet.dest.hostname_as_address())
            interface.link._transfer_packet(packet, hop)

    sc = Scenario(FRONTEND_PROGS[name], seed, replica)
    state: Dict[str, Any] = {}
    if name == "ping_pong":
        nodes = B.build_ping_pong(sc.sim)
        state = {"seen": lambda: sorted(B.seen), "clock_at_timeout": lambda: list(B.clock_at_timeout)}
    elif name == "chatter":
        nodes = B.build_chatter(sc.sim)
        state = {"heard_from": lambda: [sorted(int(a) for a in n.heard_from) for n in nodes]}
    elif name == "threads":
        log = []
        nodes = B.build_threads(sc.sim, log)
        state = {"log": lambda: list(log)}
    elif name == "family":
        graves = set()
        nodes = B.build_family(sc.sim, graves)
        state = {"graves": lambda: sorted(graves)}
    elif name == "lan":
        nodes = B.build_lan(sc.sim, 4, 16, lambda v: v, ForwardingRouter)
    elif name == "echo":
        nodes = B.build_echo(sc.sim)
        state = {"echoed": lambda: list(B.echoed), "unanswered": lambda: list(B.unanswered)}
    elif name == "daemon":
        nodes = B.build_daemon(sc.sim)
        state = {"served": lambda: list(B.served)}
    else:
        raise KeyError(name)
    for n in nodes:
        sc.node(n)
    sc.run(duration)
    final = {k: f() for k, f in state.items()}      # before result(): clearing the simulator runs the bodies' finally blocks
    arr, out = sc.result()
    out["state"] = final
    return arr, out


SCENARIOS: Dict[str, Callable] = {
    "packet_transfer": packet_transfer, "broadcast": broadcast, "thread_lifecycle": thread_lifecycle,
    "process_lifecycle": process_lifecycle, "dhcp_exchange": dhcp_exchange, "large_lan": large_lan,
}

if __name__ == "__main__":
    import json
    for name in (sys.argv[1:] or list(SCENARIOS)):
        arr, out = SCENARIOS[name]()
        print(name, json.dumps(out))
        if "-v" in sys.argv or len(arr) < 80:
            for rec in arr:
                print("   ", rec)
