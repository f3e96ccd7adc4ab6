"""
Drop-in import surface: the reference's module paths, served by this package.

``with compat.installed():`` (or ``compat.install()``) registers ``itsim``, ``itsim.simulator``, ``itsim.network.link``,
``itsim.machine.endpoint``, ``itsim.types``, ``itsim.units``, ``greensim.random`` ... in ``sys.modules`` so that a model
script -- or one of the reference's own integration tests (tests/integ/*.py) -- imports this package's classes under the
names it already uses and runs on the engine unmodified:

* ``Simulator()`` / ``Link`` / ``Endpoint`` / ``Router`` / ``Local`` / ``Relay`` / ``as_address`` / the unit constants
  are the classes of ``itsim_b200.model`` / ``itsim_b200.machine`` / ``itsim_b200.units``;
* ``node.run_proc[_in](sim, f, *args)`` takes the body *function*, as in the reference (itsim/machine/node.py:279-303);
  it is compiled by :mod:`itsim_b200.frontend` when the model is first run;
* ``sim.run(duration)`` compiles, runs replica 0 of seed ``sim.seed`` on the GPU and replays the bodies' host-side
  statements (``ledger.add(...)``, asserts on ``now()``) from the replica's trace, so that what the script inspects
  afterwards is what the reference leaves behind.  ``sim.run_replicas(n, ...)`` is the batched entry as everywhere.
* ``greensim.random`` generators are this package's declarative descriptors.

Module-level ``now()`` / ``advance()`` exist for ``from itsim.simulator import now, advance`` to succeed; outside a
compiled body they raise, like greensim's do outside a process.
"""
import contextlib
import sys
import types
from enum import IntFlag
from typing import Any, Callable, Dict, Iterator, Optional

from . import frontend, machine, model, random as _random, tags as _tags, units as _units
from .frontend import per_process  # noqa: F401

INF = float("inf")
_default_lib: Any = None         # tests inject the host emulation here


class Timeout(Exception):
    """``itsim.types.Timeout`` (itsim/types.py:220-224)."""


class ThreadKilled(Exception):
    """``itsim.machine.process_management.thread.ThreadKilled`` (thread.py:10-11)."""


class Protocol(IntFlag):
    """``itsim.types.Protocol`` (itsim/types.py:118-131), values included."""
    NONE = 0x0
    UDP = 0x1
    TCP = 0x2
    BOTH = UDP | TCP
    CLEAR = 0x00000000
    SSL = 0x80000000
    ANY = CLEAR | SSL


class Context:
    """``itsim.software.context.Context``: only ever an annotation in a compiled body."""


class Process:
    """What ``run_proc_in`` returns (itsim/machine/node.py:279-283): a handle naming the started process."""

    def __init__(self, node: Any, index: int) -> None:
        self.node, self.index = node, index


@frontend.intrinsic("now")
def now() -> float:
    raise TypeError("now() outside a compiled process body (greensim raises the same outside a process)")


@frontend.intrinsic("advance")
def advance(delay: float) -> None:
    raise TypeError("advance() outside a compiled process body")


class Simulator(model.Simulator):
    """``itsim.simulator.Simulator`` (itsim/simulator.py:9-46) on the engine."""

    def __init__(self, ts_now: float = 0.0, name: Optional[str] = None) -> None:
        super().__init__(ts_now, name)
        self.frontend = frontend.Frontend(self)
        self.seed = 0
        self.replica = 0
        self._lib = _default_lib
        self._replayed = 0
        self.caps.update(max_events=64, max_procs=128, max_sockets=64, arena_nodes=2048, trace_records=1 << 16,
                         max_threads=64, max_tasks=64, max_signals=256)
        self._default_caps = dict(self.caps)

    def compile(self) -> model.Compiled:
        procs = self._machine_procs
        for i, (node, delay, program, regs) in enumerate(procs):
            if callable(program):
                name, values = self.frontend.program(program, regs)
                procs[i] = (node, delay, name, tuple(values))
        self._size_capacities(len(procs))
        for node, _, name, _ in procs:
            if name in self.frontend.single_interface_programs and \
                    sum(1 for i in node.interfaces() if not i.address.is_loopback) != 1:
                raise frontend.CompileError(f"{name}: `for interface in context.node.interfaces()` is compiled for "
                                            f"nodes with one non-loopback interface")
        return super().compile()

    def _size_capacities(self, started: int) -> None:
        """The per-replica tables are fixed-size; the reference's are Python containers.  Capacities a script did not
        set itself grow with the number of processes the model starts (each may hold a socket, block in a ``recv`` with
        its two ``select`` helpers, and have packets queued); exceeding one still fails loudly (``ITSB_ERR_CAP_*``)."""
        wanted = dict(max_events=min(65535, started + 64), max_procs=min(8192, 2 * started + 64),
                      max_sockets=min(0x3FFF, started + 16), arena_nodes=32 * started,
                      max_threads=min(0xFFFE, started + 16), max_tasks=min(0xFFFE, started + 16),
                      max_signals=min(0x3FFF, 5 * started + 64))
        for key, value in wanted.items():
            if self.caps[key] == self._default_caps[key] and value > self.caps[key]:
                self.caps[key] = value

    def run(self, duration: float = INF, seed: Optional[int] = None) -> None:
        if self._engine is None:
            self.run_replicas(1, self.seed if seed is None else seed, first_replica_id=self.replica)
        self._engine.run(duration)
        trace = self._engine.trace(0)
        if self._engine.trace_total > len(trace):
            raise RuntimeError("trace capacity exceeded: raise sim.caps['trace_records'] to replay host-side statements")
        todo, self._replayed = trace[self._replayed:], len(trace)
        self.frontend.replay(todo)


class _BodyMixin:
    """``run_proc`` family taking body functions (node.py:279-303, 352-357)."""

    def run_proc_in(self, sim: Simulator, time: float, f: Callable, *args: Any, **kwargs: Any) -> Process:
        if kwargs:
            raise frontend.CompileError("keyword arguments to a body are not supported")
        sim._machine_procs.append((self, float(time), f, args))
        return Process(self, len(sim._machine_procs) - 1)

    def run_proc(self, sim: Simulator, f: Callable, *args: Any, **kwargs: Any) -> Process:
        return self.run_proc_in(sim, 0, f, *args, **kwargs)

    def with_proc_in(self, sim: Simulator, time: float, f: Callable, *args: Any, **kwargs: Any) -> Any:
        self.run_proc_in(sim, time, f, *args, **kwargs)
        return self

    def with_proc(self, sim: Simulator, f: Callable, *args: Any, **kwargs: Any) -> Any:
        self.run_proc(sim, f, *args, **kwargs)
        return self


class Daemon:
    """``itsim.machine.process_management.daemon.Daemon`` (daemon.py:9-27): a container for the logic run when the daemon
    is triggered."""

    def __init__(self, trigger_event: Callable[..., None]) -> None:
        self._trigger_event = trigger_event

    def trigger(self, context: Any, packet: Any, socket: Any) -> None:
        self._trigger_event(context, packet, socket)

    trigger.__itsb_inline__ = True


def _listener(daemon: Any, protocol: Any, port: int) -> Callable:
    """The body ``run_networking_daemon`` starts for one port (node.py:340-348).  The reference binds on the host before
    the run and starts ``forward_recv`` on the bound socket; here the first thread binds (outside a ``with``: a socket
    that outlives the thread), then behaves as every later one."""

    def forward_recv(context: Any, socket: Any) -> None:
        packet = socket.recv()
        context.thread.clone(forward_recv, socket)
        daemon.trigger(context, packet, socket)

    def first(context: Any) -> None:
        socket = context.node.bind(protocol, port)
        forward_recv(context, socket)

    forward_recv.__itsb_inline__ = True
    first.__name__ = "forward_recv"
    return first


class _DaemonMixin:
    """``Node.run_networking_daemon`` / ``networking_daemon`` (node.py:319-392)."""

    def run_networking_daemon(self, sim: Simulator, daemon: Any, protocol: Any, *ports: int) -> None:
        for port in ports:
            self.run_proc(sim, _listener(daemon, protocol, port))

    def networking_daemon(self, sim: Simulator, protocol: Any, *ports: int) -> Callable:
        def _decorator(server_behaviour: Any) -> Any:
            if hasattr(server_behaviour, "trigger"):
                daemon = server_behaviour if not isinstance(server_behaviour, type) else server_behaviour()
            elif callable(server_behaviour):
                daemon = Daemon(server_behaviour)
            else:
                raise TypeError("Daemon must have trigger() or be of type Callable")
            self.run_networking_daemon(sim, daemon, protocol, *ports)
            return server_behaviour
        return _decorator


class Node(_BodyMixin, _DaemonMixin, machine.Node):
    pass


class Endpoint(_BodyMixin, _DaemonMixin, machine.Endpoint):
    pass


class Router(_BodyMixin, _DaemonMixin, machine.Router):
    pass


def _module(name: str, **attrs: Any) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__itsb_compat__ = True
    return m


def _modules() -> Dict[str, types.ModuleType]:
    from ipaddress import IPv4Address as Address
    mods = {
        "itsim": _module("itsim", Tag=_tags.Tag, __path__=[]),
        "itsim.simulator": _module("itsim.simulator", Simulator=Simulator, now=now, advance=advance),
        "itsim.units": _module("itsim.units", **{k: v for k, v in vars(_units).items() if not k.startswith("_")}),
        "itsim.types": _module("itsim.types", Protocol=Protocol, Timeout=Timeout, as_address=machine.as_address,
                               Address=Address, AddressRepr=machine.AddressRepr),
        "itsim.random": _module("itsim.random", num_bytes=_random.num_bytes),
        "itsim.software": _module("itsim.software", __path__=[]),
        "itsim.software.context": _module("itsim.software.context", Context=Context),
        "itsim.machine": _module("itsim.machine", __path__=[]),
        "itsim.machine.endpoint": _module("itsim.machine.endpoint", Endpoint=Endpoint),
        "itsim.machine.node": _module("itsim.machine.node", Node=Node),
        "itsim.machine.socket": _module("itsim.machine.socket", Timeout=Timeout),
        "itsim.machine.process_management": _module("itsim.machine.process_management", __path__=[]),
        "itsim.machine.process_management.process": _module("itsim.machine.process_management.process", Process=Process),
        "itsim.machine.process_management.thread": _module("itsim.machine.process_management.thread",
                                                           ThreadKilled=ThreadKilled),
        "itsim.machine.process_management.daemon": _module("itsim.machine.process_management.daemon", Daemon=Daemon),
        "itsim.network": _module("itsim.network", __path__=[]),
        "itsim.network.link": _module("itsim.network.link", Link=machine.Link, Loopback=machine.Loopback),
        "itsim.network.router": _module("itsim.network.router", Router=Router),
        "itsim.network.route": _module("itsim.network.route", Route=machine.Route, Local=machine.Local,
                                       Relay=machine.Relay),
        "greensim": _module("greensim", now=now, advance=advance, __path__=[]),
        "greensim.random": _module("greensim.random", constant=_random.constant, uniform=_random.uniform,
                                   expo=_random.expo, normal=_random.normal, linear=_random.linear,
                                   bounded=_random.bounded, project_int=_random.project_int),
    }
    for name, mod in mods.items():          # `import itsim.network.link` then `itsim.network.link.Link`
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(mods[parent], child, mod)
    return mods


def install(lib: Any = None) -> Dict[str, Any]:
    """Registers the compatibility modules; returns what they displaced (for :func:`uninstall`)."""
    global _default_lib
    _default_lib = lib
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("itsim", "greensim")}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(_modules())
    return saved


def uninstall(saved: Dict[str, Any]) -> None:
    global _default_lib
    _default_lib = None
    for k in [k for k in sys.modules if k.split(".")[0] in ("itsim", "greensim")]:
        del sys.modules[k]
    sys.modules.update(saved)


@contextlib.contextmanager
def installed(lib: Any = None) -> Iterator[None]:
    saved = install(lib)
    try:
        yield
    finally:
        uninstall(saved)
