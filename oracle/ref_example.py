"""
ORACLE / TEST INFRASTRUCTURE ONLY -- runs the reference's own example FILES, unmodified, under the tracer.

``oracle/netsimple.py`` restates ``examples/network_simple/network_simple.py`` + ``network_simple_overrides.py``
(+ ``malware_simple.py``); this module pins that restatement: it executes the three files where they lie under
``/root/reference`` with ``runpy`` -- as ``__main__``, i.e. ``init()`` (network_simple.py:311-387) and
``sim.run(dur)`` (:391-392), resp. the module-level script of malware_simple.py:110-125 -- on the reference's own
``itsim`` package and the greensim shim, and returns the event trace, the connection records the override ``Socket``
logs, and the counters.  ``tests/test_oracle.py`` asserts that they equal ``oracle/netsimple.py``'s for several
(seed, replica) pairs.  Nothing here travels to the GPU box (``/root/reference`` does not exist there); the goldens
under ``tests/golden/`` do.

The example cannot be imported as shipped (``network_simple_overrides.py:1-6`` calls itself deprecated).
Every line that fails, and what is put in its place -- in ``sys.modules`` / as attributes of the reference's module
objects, never in its files:

=============================================  ==================================================================
failing line of the example                    what this module provides
=============================================  ==================================================================
overrides.py:23 ``from itsim.network.packet    ``PayloadDictionaryType`` does not exist anywhere in ``itsim``; ``Payload``
import Packet, Payload, PayloadDictionaryType``  is the typing alias ``Mapping[str, object]`` (itsim/types.py:15), which
(also network_simple.py:13,                    cannot be instantiated as network_simple.py:125 ``Payload({...})`` does.
malware_simple.py:10)                          Provided: an ``Enum`` with the three members the example uses (CONTENT,
                                               HOSTNAME, ADDRESS) and a ``Payload`` class holding the dict as
                                               ``entries`` / ``_entries`` (the two spellings the example reads:
                                               network_simple.py:140, overrides.py:159)
overrides.py:24 ``from itsim.machine import    ``itsim/machine/__init__.py`` defines only the abstract ``_Node`` /
Node as ParentNode, Socket as ParentSocket``   ``_Socket``.  ``Node`` := the reference's real ``itsim.machine.node.Node``
                                               (subclassed only to add ``_as_source_bind``, below).  ``Socket`` := a bare
                                               base: overrides.py:68 calls ``super().__init__(src, node)``, which the
                                               shipped ``Socket.__init__(protocol, port, node, pid)`` (socket.py:28)
                                               rejects, and the shipped ``__del__`` / ``close`` would index the override
                                               node's ``_sockets`` by port (it is keyed by Location there)
overrides.py:339 ``self._as_source_bind(lb)``  defined nowhere in the reference.  Provided, from its three call shapes
                                               (``open_socket(68)``, ``open_socket()``, ``open_socket(Location)``):
                                               None -> Location(default address, 0); int -> Location(default address,
                                               port); Location -> itself
=============================================  ==================================================================

Canonicalisations, none of which touches the example's text:

* every ``greensim.random`` generator the example builds with ``rng=None`` (all of them: network_simple.py:101-103,
  128,171,249-250; overrides.py:451-452; malware_simple.py:56,63,114) draws from ONE Philox stream per replica --
  ``greensim.random.set_default_rng`` -- instead of the process-global Mersenne Twister;
* ``greensim.Simulator`` is, for the duration of the run, a subclass that attaches the tracer when the example
  constructs it (network_simple.py:331) -- pure observation;
* processes are labelled for the tracer from what they run (``Process._body``: function name -> program id, first
  argument / bound ``self`` / closure -> node index), the labels ``oracle/netsimple.py`` attaches by hand;
* the example's loggers get handlers before it runs, so that it does not print ~1e5 lines; the ``JSON-OUT`` records of
  ``Socket.log_pair`` (overrides.py:165-177) are captured and parsed: they are the connection records.
"""
import inspect
import json
import logging
import os
import runpy
import sys
from enum import Enum
from ipaddress import ip_address
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ITSIM_REFERENCE", "/root/reference")
EXAMPLE_DIR = os.path.join(REFERENCE, "examples", "network_simple")
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import evtrace as tr  # noqa: E402
from philox import PhiloxRandom  # noqa: E402

# program ids: the numbering oracle/netsimple.py and itsim_b200/netsimple.py share
PROGRAMS = {
    "dhcp_serve": 1, "workstation_blinking": 2, "client_activity": 3, "mdns_daemon": 4, "llmnr_daemon": 5,
    "_query_mdns": 6, "_query_llmnr": 7, "_wait_for_timeout": 8, "transmission": 9, "_receive": 10,
    "call_home": 11, "no_good": 12, "sudo_transmit": 13, "exfiltrate": 14, "route": 15, "command_and_control": 16,
}
EXC_COMPLETE = 1


def available() -> bool:
    return os.path.isfile(os.path.join(EXAMPLE_DIR, "network_simple.py"))


# ---- the names the shipped itsim lacks ---------------------------------------------------------------------------
class PayloadDictionaryType(Enum):
    """Keys of the example's payload dictionaries (network_simple.py:125,197-219; overrides.py:159)."""
    CONTENT = "content"
    HOSTNAME = "hostname"
    ADDRESS = "address"


class Payload:
    """What ``Payload({...})`` must be for network_simple.py:140 (``.entries``) and overrides.py:159 (``._entries``)."""

    def __init__(self, entries: Optional[Dict[Any, Any]] = None) -> None:
        self._entries = dict(entries or {})

    @property
    def entries(self) -> Dict[Any, Any]:
        return self._entries

    def __repr__(self) -> str:
        return f"Payload({self._entries!r})"


class _BareSocket:
    """Parent of the override Socket (overrides.py:65-68): accepts ``(src, node)``, keeps nothing."""

    def __init__(self, src: Any = None, node: Any = None) -> None:
        pass


def _install_missing_names() -> None:
    """Imports the reference's itsim (unmodified, on the greensim shim) and adds the names above to its modules."""
    if REFERENCE not in sys.path:
        sys.path.insert(1, REFERENCE)
    import greensim
    assert getattr(greensim, "GREENSIM_TAG_ATTRIBUTE", None), "oracle/greensim shim expected"
    import itsim.machine
    import itsim.machine.node
    import itsim.network.packet
    from itsim.network.location import Location

    itsim.network.packet.Payload = Payload
    itsim.network.packet.PayloadDictionaryType = PayloadDictionaryType

    class Node(itsim.machine.node.Node):
        """The reference's shipped Node + the one method overrides.py:339 calls and nothing defines."""

        def _as_source_bind(self, lb: Any = None) -> Location:
            if isinstance(lb, Location):
                return lb
            return Location(self.address_default, lb or 0)

    itsim.machine.Node = Node
    itsim.machine.Socket = _BareSocket


# ---- observation -------------------------------------------------------------------------------------------------
class _Capture(logging.Handler):
    def __init__(self) -> None:
        super().__init__(level=logging.INFO)
        self.lines: List[str] = []

    def emit(self, record: logging.LogRecord) -> None:
        self.lines.append(record.getMessage())


class Observer:
    """Labels the example's processes for the tracer and counts what oracle/netsimple.py counts by hand."""

    def __init__(self) -> None:
        self.sim: Any = None
        self.tracer: Optional[tr.Tracer] = None
        self._bases: Dict[int, int] = {}      # id(Network) -> index of its first node
        self._nets: List[Any] = []
        self._next_base = 0

    # node numbering: networks in order of first use, nodes in link order inside each (the order Network.transmit
    # walks them: overrides.py:282); a second attachment of a machine (malware_simple.py:119) is numbered on its own
    def _base(self, net: Any) -> int:
        if id(net) not in self._bases:
            self._bases[id(net)] = self._next_base
            self._nets.append(net)
            self._next_base += len(net._nodes)
        return self._bases[id(net)]

    def _index(self, node: Any, net: Any) -> int:
        base = self._base(net)
        for i, member in enumerate(net._nodes.values()):
            if member is node:
                return base + i
        raise LookupError("node not on this network")

    def _network_of(self, node: Any, address: Any) -> Any:
        for link in node._networks.values():
            if address == link.address or address == link.network.address_broadcast:
                return link.network
        return next(iter(node._networks.values())).network

    def attach(self, sim: Any) -> None:
        self.sim = sim
        self.tracer = tr.Tracer(sim, keep_records=True)
        self.tracer.resolver = self.resolve
        self.tracer.exc_code = lambda exc: EXC_COMPLETE if type(exc).__name__ == "_Complete" else 0
        self.tracer.deliver_aux = self.deliver_aux

    def _args_of(self, proc: Any) -> tuple:
        """Arguments of a process that has not started: those of the event that is about to switch into it; a process
        first seen through a throw (never the case for the example's bodies) has none."""
        ev = self.tracer.current_event
        if ev is not None and getattr(ev.fn, "__self__", None) is proc and ev.fn.__name__ == "switch":
            return ev.args
        return proc._inbox_args

    def resolve(self, proc: Any) -> tr.Ident:
        body = proc._body
        name = getattr(body, "__name__", "")
        prog = PROGRAMS.get(name, 0)
        if name == "transmission":          # the closure of Network.transmit (overrides.py:286-290)
            cells = inspect.getclosurevars(body).nonlocals
            packet, net = cells["packet"], cells["self"]
            self._base(net)
            sender = net._nodes[packet.source.hostname_as_address()]
            return tr.Ident(tr.ROLE_TX, prog, self._index(sender, net), packet.byte_size)
        if name == "_receive":              # node._receive(packet), one process per recipient (overrides.py:290)
            node, packet = body.__self__, self._args_of(proc)[0]
            return tr.Ident(tr.ROLE_DELIVER, prog, self._index(node, self._network_of(node, packet.dest.hostname)))
        if name == "_wait_for_timeout":     # Workstation.timeout's helper (network_simple.py:84-98)
            ws = body.__self__
        else:                               # installed software and what it adds: the endpoint is the first argument
            args = self._args_of(proc)       # (ws, ...) -- or (logger, ws, ...) for the two query processes
            ws = args[0] if hasattr(args[0], "_networks") else args[1]
        return tr.Ident(tr.ROLE_MODEL, prog, self._index(ws, ws._network))

    @staticmethod
    def deliver_aux(proc: Any, ev: Any) -> int:
        node, packet = proc._body.__self__, ev.args[0]
        accepted = 1 if packet.dest in node._sockets else 0
        return (packet.byte_size & 0x7FFFFFFF) | (accepted << 31)


def _connection_tuple(line: str, index_of_ip: Dict[str, int]) -> Tuple:
    """One JSON-OUT record of Socket.log_pair (overrides.py:165-177) in the field order of
    ``oracle/netsimple.Socket._log_pair``."""
    d = json.loads(line)

    def ip(v: Any) -> int:
        return 0 if v in (0, "0") else int(ip_address(v))

    return (d["End_Time"], ip(d["Local_IP"]), int(d["Local_Port"]), 1 if d["Direction"] == "Inbound" else 0,
            ip(d["Remote_IP"]), int(d["Remote_Port"]), int(d["Sent_Bytes"]), int(d["Received_Bytes"]), d["Start_Time"])


def run_example(script: str, seed: int, replica: int, duration: float) -> Dict[str, Any]:
    """Executes ``examples/network_simple/<script>`` unmodified for ``duration`` simulated seconds on Philox stream
    (seed, replica).  Returns {"trace", "summary", "connections", "draws", "now"}."""
    assert available(), f"{EXAMPLE_DIR} not found"
    _install_missing_names()
    import greensim
    import greensim.random as gsrandom

    observer = Observer()
    stock_simulator = greensim.Simulator

    class TracedSimulator(stock_simulator):          # what network_simple.py:331 `Simulator()` builds during this run
        def __init__(self, *args: Any, **kwargs: Any) -> None:
            super().__init__(*args, **kwargs)
            observer.attach(self)

    # loggers: quiet, except the connection records
    capture = _Capture()
    quiet = logging.NullHandler()
    names = ["blinking", "dhcp_client", "dhcp_server", "mdns_responder", "llmnr_responder", "client_activity", "main",
             "network_simple", "__main__"]
    saved_levels = {}
    for name in names:
        lg = logging.getLogger(name)
        saved_levels[name] = (lg.level, list(lg.handlers), lg.propagate)
        lg.handlers = [quiet]
        lg.setLevel(logging.CRITICAL)
        lg.propagate = False
    conn_logger = logging.getLogger("network_simple_overrides")
    saved_conn = (conn_logger.level, list(conn_logger.handlers), conn_logger.propagate)
    conn_logger.handlers = [capture]
    conn_logger.setLevel(logging.INFO)
    conn_logger.propagate = False
    root_level = logging.getLogger().level

    rng = PhiloxRandom(seed, replica)
    saved_argv, saved_path = sys.argv, list(sys.path)
    for mod in ("network_simple", "network_simple_overrides", "malware_simple"):
        sys.modules.pop(mod, None)
    sys.argv = [script, "-d", repr(float(duration)), "-o"]
    sys.path.insert(0, EXAMPLE_DIR)
    greensim.Simulator = TracedSimulator
    gsrandom.set_default_rng(rng)
    try:
        runpy.run_path(os.path.join(EXAMPLE_DIR, script), run_name="__main__")
    finally:
        gsrandom.set_default_rng(None)
        greensim.Simulator = stock_simulator
        sys.argv, sys.path[:] = saved_argv, saved_path
        for name, (level, handlers, propagate) in saved_levels.items():
            lg = logging.getLogger(name)
            lg.handlers, lg.propagate = handlers, propagate
            lg.setLevel(level)
        conn_logger.handlers, conn_logger.propagate = saved_conn[1], saved_conn[2]
        conn_logger.setLevel(saved_conn[0])
        logging.getLogger().setLevel(root_level)
        for mod in ("network_simple", "network_simple_overrides", "malware_simple"):
            sys.modules.pop(mod, None)
    tracer = observer.tracer
    assert tracer is not None, "the example never built a Simulator"
    out = {
        "trace": tracer.as_array(),
        "summary": tracer.summary(),
        "connections": [_connection_tuple(line, {}) for line in capture.lines],
        "draws": rng.draws,
        "now": observer.sim.now(),
    }
    observer.sim._clear()
    return out


def run_network_simple(seed: int, replica: int, duration: float) -> Dict[str, Any]:
    return run_example("network_simple.py", seed, replica, duration)


def run_malware_simple(seed: int, replica: int, duration: float) -> Dict[str, Any]:
    return run_example("malware_simple.py", seed, replica, duration)


def compare_with_restatement(script: str, seed: int, replica: int, duration: float) -> Dict[str, Any]:
    """Runs the reference's file and oracle/netsimple.py on the same stream; returns both results."""
    import netsimple
    ref = run_example(script, seed, replica, duration)
    model = (netsimple.MalwareModel if script == "malware_simple.py" else netsimple.Model)(seed, replica)
    model.run(duration)
    ours = {
        "trace": model.tracer.as_array(),
        "summary": model.tracer.summary(),
        # oracle record: (t_end, node, local_ip, local_port, inbound, remote_ip, remote_port, sent, received, start)
        "connections": [(c[0], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9]) for c in model.connections],
        "draws": model.rng.draws,
        "now": model.sim.now(),
    }
    model.close()
    return {"reference": ref, "restatement": ours}


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "network_simple.py"
    dur = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
    both = compare_with_restatement(which, 0, 0, dur)
    a, b = both["reference"], both["restatement"]
    same = len(a["trace"]) == len(b["trace"]) and all(
        np.array_equal(a["trace"][f], b["trace"][f]) for f in ("t", "kind", "prog", "node", "aux"))
    print(json.dumps({"script": which, "duration": dur, "events_reference": a["summary"]["events"],
                      "events_restatement": b["summary"]["events"], "trace_identical": bool(same),
                      "connections_identical": a["connections"] == b["connections"],
                      "draws": [a["draws"], b["draws"]]}))
