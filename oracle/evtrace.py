"""
ORACLE / TEST INFRASTRUCTURE ONLY.

The shared definition of a "simulated event"  and the tracer that extracts it from the
greensim shim.  One trace record = one *live* heap pop that starts or resumes model-visible work:

====  ===========  ==============================================================================
kind  name         what popped
====  ===========  ==============================================================================
1     PROC_START   first switch into a model process body (``add`` / ``add_in`` / ``install``)
2     TIMER        return from ``advance`` in a model process
3     TX_BEGIN     start of a transmission process (override path ``Network.transmit``)
4     TX_END       its ``advance(latency + len/bandwidth)`` elapsed; deliveries are spawned
5     DELIVER      a recipient's ``_receive`` / ``_receive_packet`` runs (one per recipient, drops included)
6     WAKE         a process resumes from a ``Signal``/``Queue`` wait (recv, wake-up, join, wait)
7     THROW        an exception is delivered into a live model process (timeout, kill, ...)
8     STOP         the run-horizon event
====  ===========  ==============================================================================

Not events: hops of the engine's own helper processes (``select`` / balk helpers and the CleanUp / CancelBalk
interrupts sent to them) and switches/throws into finished processes ("ghost" hops).  They are counted
separately (``hidden``, ``ghost``) so raw heap pops can be reported too.

Record layout (identical to ``itsb_trace_rec`` in include/itsb.h): ``t f64, kind u8, prog u8, node u16, aux u32``.
"""
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

PROC_START, TIMER, TX_BEGIN, TX_END, DELIVER, WAKE, THROW, STOP = 1, 2, 3, 4, 5, 6, 7, 8
KIND_NAMES = {1: "PROC_START", 2: "TIMER", 3: "TX_BEGIN", 4: "TX_END", 5: "DELIVER", 6: "WAKE", 7: "THROW", 8: "STOP"}
NUM_KINDS = 9

TRACE_DTYPE = np.dtype([("t", "<f8"), ("kind", "u1"), ("prog", "u1"), ("node", "<u2"), ("aux", "<u4")])

# process roles on the special transmission / delivery processes
ROLE_MODEL, ROLE_TX, ROLE_DELIVER = 0, 1, 2


class Ident:
    """What the tracer knows about one greensim process."""
    __slots__ = ("role", "prog", "node", "aux", "blocked_in")

    def __init__(self, role: int, prog: int, node: int, aux: int = 0) -> None:
        self.role = role
        self.prog = prog
        self.node = node
        self.aux = aux
        self.blocked_in = None  # "advance" | "wait": what the process is parked in (set by the shim hooks)


class Tracer:
    """Attach to a shim ``Simulator``; model code labels its processes through :meth:`label`."""

    def __init__(self, sim: Any, keep_records: bool = True, limit: Optional[int] = None) -> None:
        self.sim = sim
        self.keep = keep_records
        self.limit = limit
        self.records: List[Tuple[float, int, int, int, int]] = []
        self.counts = [0] * NUM_KINDS
        self.hidden = 0
        self.ghost = 0
        self.exc_code: Callable[[BaseException], int] = lambda exc: 0
        self.deliver_aux: Callable[[Any, Any], int] = lambda proc, ev: 0
        self.resolver: Optional[Callable[[Any], "Ident"]] = None   # labels processes the model code did not label
        self.current_event: Any = None
        sim._on_pop = self._on_pop
        self._stop_fn = sim.stop

    # model side -------------------------------------------------------------------------------------
    @staticmethod
    def label(process: Any, role: int, prog: int, node: int, aux: int = 0) -> Any:
        process.ident = Ident(role, prog, node, aux)
        return process

    # shim side --------------------------------------------------------------------------------------
    def _emit(self, kind: int, prog: int, node: int, aux: int) -> None:
        self.counts[kind] += 1
        if self.keep and (self.limit is None or len(self.records) < self.limit):
            self.records.append((self.sim.now(), kind, prog, node, aux & 0xFFFFFFFF))

    def _on_pop(self, ev: Any) -> None:
        fn = ev.fn
        target = getattr(fn, "__self__", None)
        if target is self.sim:
            if fn.__name__ == "stop":
                self._emit(STOP, 0, 0, 0)
            return
        name = fn.__name__
        if target is None or name not in ("switch", "throw"):
            return
        proc = target
        if proc.dead:
            self.ghost += 1
            return
        if proc.hidden:
            self.hidden += 1
            return
        ident: Optional[Ident] = getattr(proc, "ident", None)
        self.current_event = ev         # what is being popped (a resolver reads a new process's arguments off it)
        if ident is None:
            ident = self.resolver(proc) if self.resolver is not None else Ident(ROLE_MODEL, 0, 0xFFFF)
            proc.ident = ident
        if name == "throw":
            exc = ev.args[0]
            if getattr(exc, "hidden", False):
                self.hidden += 1
                return
            self._emit(THROW, ident.prog, ident.node, self.exc_code(exc))
            return
        if not proc._started:
            if ident.role == ROLE_TX:
                self._emit(TX_BEGIN, ident.prog, ident.node, ident.aux)
            elif ident.role == ROLE_DELIVER:
                self._emit(DELIVER, ident.prog, ident.node, self.deliver_aux(proc, ev))
            else:
                self._emit(PROC_START, ident.prog, ident.node, ident.aux)
            return
        blocked_in = getattr(proc, "_blocked_in", None)
        if blocked_in == "advance":
            if ident.role == ROLE_TX:
                self._emit(TX_END, ident.prog, ident.node, ident.aux)
            else:
                self._emit(TIMER, ident.prog, ident.node, 0)
        else:
            self._emit(WAKE, ident.prog, ident.node, 0)

    # results ----------------------------------------------------------------------------------------
    @property
    def num_events(self) -> int:
        return sum(self.counts)

    def as_array(self) -> np.ndarray:
        out = np.zeros(len(self.records), dtype=TRACE_DTYPE)
        if self.records:
            cols = list(zip(*self.records))
            out["t"], out["kind"], out["prog"], out["node"], out["aux"] = cols
        return out

    def summary(self) -> Dict[str, Any]:
        return {
            "events": self.num_events,
            "counts": {KIND_NAMES[k]: self.counts[k] for k in range(1, NUM_KINDS)},
            "hidden_hops": self.hidden,
            "ghost_hops": self.ghost,
            "raw_pops": self.sim.num_pops,
        }


def trace_digest(arr: np.ndarray) -> str:
    """Order-sensitive digest of the integer columns of a trace (timestamps are compared with a tolerance)."""
    import hashlib
    packed = np.zeros(len(arr), dtype=np.dtype([("kind", "u1"), ("prog", "u1"), ("node", "<u2"), ("aux", "<u4")]))
    for name in ("kind", "prog", "node", "aux"):
        packed[name] = arr[name]
    return hashlib.sha256(packed.tobytes()).hexdigest()
