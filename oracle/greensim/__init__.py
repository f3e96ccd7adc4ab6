"""
ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package ``itsim_b200``.

A ``greensim``-compatible discrete-event kernel, written from the published behaviour of
ElementAI/greensim (the engine the reference ITsim delegates to: ``itsim/simulator.py:4,9,80-83``,
``itsim/machine/socket.py:35-36,134-135``, ``itsim/network/link.py:86``,
``itsim/machine/process_management/thread.py:47,98``).  greensim itself is *not* in
``/root/reference`` (unpinned dependency: ``Pipfile:7`` ``greensim = "*"``, ``setup.py:24``) and neither it
nor ``greenlet`` is installed in this image, so this file restates the contract listed in SURVEY.md
Appendix A so that the reference's *unmodified* ``itsim`` package runs on top of it.  It is pinned by
running the reference's own test-suite against it (``oracle/run_reference_tests.py``).

Semantics restated here (what the B200 engine must reproduce event for event):

* event heap keyed ``(timestamp, insertion counter)``; ``_schedule(delay, fn, *args)`` pushes
  ``(now + delay, counter++)``; negative delay -> ``ValueError``.
* ``add/add_in/add_at`` create a :class:`Process` and schedule its first switch.
* ``run(duration)`` schedules a stop event first (taking the counter value at call time), pops
  until stopped or empty, and neutralises the stop event if it was not what ended the run.
* ``advance(d)`` schedules the caller's wake-up and yields; an :class:`Interrupt` thrown in cancels
  that wake-up and propagates; any *other* exception thrown in leaves the wake-up in the heap, where
  it later no-ops on the finished process (a "ghost" hop).
* ``Process.resume()`` = ``_schedule(0, switch)``; ``Process.interrupt(exc)`` = ``_schedule(0, throw, exc)``.
  Switching or throwing into a finished process is a no-op.
* ``Queue.join(timeout)``: FIFO by join counter; a timeout is a helper ("balk") process started with
  ``add`` that advances then interrupts the waiter with :class:`Timeout`; the helper is interrupted
  (``CancelBalk``) when the waiter leaves.  ``Signal``: level triggered, constructed *on*;
  ``turn_on`` resumes every waiter in FIFO order; ``wait`` loops ``while not on: queue.join``.
* ``select(*signals, timeout)``: one helper process per signal, a private signal for the caller,
  helpers interrupted (``CleanUp``) on exit.

Coroutines: greenlet is unavailable, so each simulated process body runs on a pooled OS thread with a
strict baton hand-off (exactly one runnable thread at any time).  Slow, but exact.
"""
import heapq
import threading
import weakref
from math import inf
from typing import Any, Callable, List, Optional
from uuid import uuid4
import _thread

from .tags import Tags, TaggedObject  # noqa: F401

__all__ = [
    "Simulator", "Process", "Interrupt", "Timeout", "Signal", "Queue", "select", "add", "add_in",
    "add_at", "advance", "pause", "now", "stop", "local", "tagged", "happens", "Named", "GREENSIM_TAG_ATTRIBUTE",
]

GREENSIM_TAG_ATTRIBUTE = "_greensim_tags"

threading.stack_size(512 * 1024)


class Interrupt(Exception):
    """Raised on a process by :meth:`Process.interrupt`; ends the process silently if it is not caught."""
    pass


class Timeout(Interrupt):
    """Interrupt delivered to a process that waited too long in a :class:`Queue`."""
    pass


class _Unwind(BaseException):
    """Internal: unwinds a parked body when its simulator is shut down (GreenletExit analogue)."""
    pass


class Named:

    def __init__(self, name: Optional[str] = None) -> None:
        self._name = name if name is not None else str(uuid4())

    @property
    def name(self) -> str:
        return self._name


def tagged(*tags: Tags) -> Callable:
    """Decorator: processes built from the decorated function carry ``tags``."""
    def decorate(fn: Callable) -> Callable:
        setattr(fn, GREENSIM_TAG_ATTRIBUTE, tuple(tags))
        return fn
    return decorate


# ------------------------------------------------------------------------------------------------------------
# Coroutine plumbing: pooled worker threads with baton hand-off.
# ------------------------------------------------------------------------------------------------------------

_tls = threading.local()


class _Worker:
    """An OS thread that runs successive process bodies.  The simulator side and the worker side each own
    one lock used as a binary semaphore."""

    _idle: List["_Worker"] = []
    _idle_guard = threading.Lock()

    def __init__(self) -> None:
        self.go = _thread.allocate_lock()
        self.go.acquire()
        self.job: Optional["Process"] = None
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    @classmethod
    def obtain(cls) -> "_Worker":
        with cls._idle_guard:
            if cls._idle:
                return cls._idle.pop()
        return cls()

    def _loop(self) -> None:
        while True:
            self.go.acquire()
            proc = self.job
            assert proc is not None
            _tls.process = proc
            proc._execute_body()
            _tls.process = None
            self.job = None
            back = proc._back
            with _Worker._idle_guard:
                _Worker._idle.append(self)
            back.release()  # hand the baton back to the simulator side for the last time.


class _Event:
    __slots__ = ("timestamp", "identifier", "fn", "args", "kwargs", "cancelled")

    def __init__(self, timestamp: float, identifier: int, fn: Callable, args: tuple, kwargs: dict) -> None:
        self.timestamp = timestamp
        self.identifier = identifier
        self.fn = fn
        self.args = args
        self.kwargs = kwargs
        self.cancelled = False

    def __lt__(self, other: "_Event") -> bool:
        return (self.timestamp, self.identifier) < (other.timestamp, other.identifier)


class Simulator(Named):
    """Event loop.  See module docstring for the contract."""

    def __init__(self, ts_now: float = 0.0, name: Optional[str] = None) -> None:
        super().__init__(name)
        self._ts_now = ts_now
        self._events: List[_Event] = []
        self._is_running = False
        self._counter = 0
        self._processes_alive: "weakref.WeakSet[Process]" = weakref.WeakSet()
        # Observation hooks used only by the oracle's tracer (oracle/evtrace.py); None in plain use.
        self._on_pop: Optional[Callable[[_Event], None]] = None
        self.num_pops = 0

    # -- time & state ------------------------------------------------------------------------------
    def now(self) -> float:
        return self._ts_now

    @property
    def is_running(self) -> bool:
        return self._is_running

    def events(self):
        return ((ev.timestamp, ev.fn, ev.args, ev.kwargs) for ev in sorted(self._events) if not ev.cancelled)

    # -- scheduling --------------------------------------------------------------------------------
    def _schedule(self, delay: float, fn: Callable, *args: Any, **kwargs: Any) -> int:
        if delay < 0.0:
            raise ValueError("Delay must be positive.")
        identifier = self._counter
        self._counter += 1
        heapq.heappush(self._events, _Event(self._ts_now + delay, identifier, fn, args, kwargs))
        return identifier

    def _cancel(self, identifier: int) -> None:
        for ev in self._events:
            if ev.identifier == identifier:
                ev.cancelled = True
                break
        self._events = [ev for ev in self._events if not ev.cancelled]
        heapq.heapify(self._events)

    def add(self, fn_process: Callable, *args: Any, **kwargs: Any) -> "Process":
        return self.add_in(0.0, fn_process, *args, **kwargs)

    def add_in(self, delay: float, fn_process: Callable, *args: Any, **kwargs: Any) -> "Process":
        process = Process(self, fn_process, None)
        self._schedule(delay, process.switch, *args, **kwargs)
        return process

    def add_at(self, moment: float, fn_process: Callable, *args: Any, **kwargs: Any) -> "Process":
        delay = moment - self.now()
        if delay < 0.0:
            raise ValueError(f"The given moment to start the process ({moment:f}) is in the past.")
        return self.add_in(delay, fn_process, *args, **kwargs)

    # -- running -----------------------------------------------------------------------------------
    def run(self, duration: float = inf) -> None:
        id_stop = None
        if duration != inf:
            id_stop = self._schedule(duration, self.stop)

        self._is_running = True
        stopped_by_own_event = False
        while self._is_running and self._events:
            ev = heapq.heappop(self._events)
            self._ts_now = ev.timestamp
            self.num_pops += 1
            if self._on_pop is not None:
                self._on_pop(ev)
            if id_stop is not None and ev.identifier == id_stop:
                stopped_by_own_event = True
            ev.fn(*ev.args, **ev.kwargs)

        if id_stop is not None and not stopped_by_own_event:
            self._cancel(id_stop)
        self.stop()

    def step(self) -> None:
        if self._events:
            ev = heapq.heappop(self._events)
            self._ts_now = ev.timestamp
            self.num_pops += 1
            if self._on_pop is not None:
                self._on_pop(ev)
            ev.fn(*ev.args, **ev.kwargs)

    def stop(self) -> None:
        self._is_running = False

    def _clear(self) -> None:
        """Discards pending events and unwinds every parked process body (oracle housekeeping)."""
        self._events = []
        for process in list(self._processes_alive):
            process._unwind()


class _LocalData:
    pass


class Process(TaggedObject):
    """A simulated process: a coroutine whose body may block in ``advance``/``pause`` from arbitrarily nested
    plain calls."""

    def __init__(self, sim: Simulator, body: Callable, parent: Any = None) -> None:
        tags = list(getattr(body, GREENSIM_TAG_ATTRIBUTE, ()))
        creator = getattr(_tls, "process", None)
        if creator is not None:
            tags.extend(creator.iter_tags())
        super().__init__(*tags)
        self._sim = sim
        self.rsim = weakref.ref(sim)
        self._body = body
        self._started = False
        self._finished = False
        self._worker: Optional[_Worker] = None
        self._back = _thread.allocate_lock()  # simulator side parks here while the body runs
        self._back.acquire()
        self._inbox_args: tuple = ()
        self._inbox_kwargs: dict = {}
        self._inbox_exc: Optional[BaseException] = None
        self._failure: Optional[BaseException] = None
        self.local = _LocalData()
        self.local.name = str(uuid4())
        # True for the helper processes this kernel spawns itself (select / balk); the oracle's tracer does not
        # count their hops as model events.
        self.hidden = False
        self._blocked_in: Optional[str] = None  # "advance" | "wait" while parked (read by the tracer)

    # -- identity ----------------------------------------------------------------------------------
    @staticmethod
    def current() -> "Process":
        proc = getattr(_tls, "process", None)
        if proc is None:
            raise TypeError("Current thread of control does not correspond to a Process instance.")
        return proc

    @property
    def dead(self) -> bool:
        return self._finished

    def __bool__(self) -> bool:  # greenlet truthiness: started and not finished
        return self._started and not self._finished

    # -- simulator side ----------------------------------------------------------------------------
    def switch(self, *args: Any, **kwargs: Any) -> None:
        if self._finished:
            return  # switching into a finished coroutine: no-op hop
        if not self._started:
            self._started = True
            self._inbox_args, self._inbox_kwargs = args, kwargs
            self._sim._processes_alive.add(self)
            self._worker = _Worker.obtain()
            self._worker.job = self
        self._transfer()

    def throw(self, exc: BaseException) -> None:
        if self._finished:
            return  # throwing into a finished coroutine: no-op hop
        if not self._started:
            # A never-started coroutine dies on the spot; an Interrupt is swallowed like in a body.
            self._started = True
            self._finished = True
            if not isinstance(exc, (Interrupt, _Unwind)):
                raise exc
            return
        self._inbox_exc = exc
        self._transfer()

    def _transfer(self) -> None:
        worker = self._worker
        assert worker is not None
        caller = getattr(_tls, "process", None)
        assert caller is None, "only the simulator's thread of control may switch into a process"
        worker.go.release()
        self._back.acquire()
        if self._failure is not None:
            failure, self._failure = self._failure, None
            raise failure

    def _unwind(self) -> None:
        if self._started and not self._finished:
            try:
                self.throw(_Unwind())
            except BaseException:
                pass

    # -- worker side -------------------------------------------------------------------------------
    def _execute_body(self) -> None:
        try:
            self._body(*self._inbox_args, **self._inbox_kwargs)
        except Interrupt:
            pass
        except _Unwind:
            pass
        except BaseException as err:  # propagates out of Simulator.run, like an error in a greenlet body
            self._failure = err
        finally:
            self._finished = True
            self._inbox_args, self._inbox_kwargs = (), {}
            self._sim._processes_alive.discard(self)

    def _yield_to_simulator(self, reason: str = "wait") -> None:
        """Called on the worker: give the baton back and wait for the next switch/throw."""
        worker = self._worker
        assert worker is not None
        self._blocked_in = reason
        self._back.release()
        worker.go.acquire()
        _tls.process = self
        exc, self._inbox_exc = self._inbox_exc, None
        if exc is not None:
            raise exc

    # -- scheduling helpers ------------------------------------------------------------------------
    def resume(self) -> None:
        self._sim._schedule(0.0, self.switch)

    def interrupt(self, inter: Optional[Interrupt] = None) -> None:
        if inter is None:
            inter = Interrupt()
        self._sim._schedule(0.0, self.throw, inter)


class _Local:
    """``greensim.local`` -- attribute proxy to the current process's local data."""

    def __getattr__(self, name: str) -> Any:
        return getattr(Process.current().local, name)

    def __setattr__(self, name: str, value: Any) -> None:
        setattr(Process.current().local, name, value)


local = _Local()


def pause() -> None:
    Process.current()._yield_to_simulator("wait")


def advance(delay: float) -> None:
    proc = Process.current()
    sim = proc._sim
    id_wakeup = sim._schedule(delay, proc.switch)
    try:
        proc._yield_to_simulator("advance")
    except Interrupt:
        sim._cancel(id_wakeup)
        raise


def now() -> float:
    return Process.current()._sim.now()


def add(fn_process: Callable, *args: Any, **kwargs: Any) -> Process:
    return Process.current()._sim.add(fn_process, *args, **kwargs)


def add_in(delay: float, fn_process: Callable, *args: Any, **kwargs: Any) -> Process:
    return Process.current()._sim.add_in(delay, fn_process, *args, **kwargs)


def add_at(moment: float, fn_process: Callable, *args: Any, **kwargs: Any) -> Process:
    return Process.current()._sim.add_at(moment, fn_process, *args, **kwargs)


def stop() -> None:
    Process.current()._sim.stop()


def happens(intervals, name: Optional[str] = None) -> Callable:
    """Decorator: the decorated function becomes an event spawned after each interval drawn from ``intervals``."""
    def hook(event: Callable) -> Callable:
        def make_happen(*args_event: Any, **kwargs_event: Any) -> None:
            if name is not None:
                local.name = name
            for interval in intervals:
                advance(interval)
                add(event, *args_event, **kwargs_event)
        return make_happen
    return hook


class Queue(Named):
    """Waiting line of processes; FIFO unless an order-token function says otherwise."""

    def __init__(self, get_order_token: Optional[Callable[[int], int]] = None, name: Optional[str] = None) -> None:
        super().__init__(name)
        self._waiting: List[Any] = []
        self._counter = 0
        self._get_order_token = get_order_token or (lambda counter: counter)

    def is_empty(self) -> bool:
        return len(self._waiting) == 0

    def __len__(self) -> int:
        return len(self._waiting)

    def peek(self) -> Process:
        return self._waiting[0][2]

    def join(self, timeout: Optional[float] = None) -> None:
        class CancelBalk(Interrupt):
            hidden = True

        me = Process.current()
        self._counter += 1
        heapq.heappush(self._waiting, (self._get_order_token(self._counter), self._counter, me))

        balker: List[Optional[Process]] = [None]
        if timeout is not None:
            def balk(proc: Process) -> None:
                try:
                    advance(timeout)
                    proc.interrupt(Timeout())
                except CancelBalk:
                    pass
                finally:
                    balker[0] = None
            balker[0] = add(balk, me)
            balker[0].hidden = True

        try:
            pause()
        except Interrupt:
            self._waiting = [entry for entry in self._waiting if entry[2] is not me]
            heapq.heapify(self._waiting)
            raise
        finally:
            if balker[0] is not None:
                balker[0].interrupt(CancelBalk())

    def pop(self) -> None:
        if self._waiting:
            _, _, process = heapq.heappop(self._waiting)
            process.resume()


class Signal(Named):
    """Level-triggered gate: processes wait while it is off; turning it on releases all of them, in order."""

    def __init__(self, get_order_token: Optional[Callable[[int], int]] = None, name: Optional[str] = None) -> None:
        super().__init__(name)
        self._is_on = True
        self._queue = Queue(get_order_token)

    @property
    def is_on(self) -> bool:
        return self._is_on

    def turn_on(self) -> "Signal":
        self._is_on = True
        while not self._queue.is_empty():
            self._queue.pop()
        return self

    def turn_off(self) -> "Signal":
        self._is_on = False
        return self

    def wait(self, timeout: Optional[float] = None) -> None:
        while not self._is_on:
            self._queue.join(timeout)


def select(*signals: Signal, **kwargs: Any) -> List[Signal]:
    """Blocks until one of the signals is on, or ``timeout`` (keyword) elapses -> :class:`Timeout`."""
    class CleanUp(Interrupt):
        hidden = True

    timeout = kwargs.get("timeout", None)
    if not isinstance(timeout, (float, int, type(None))):
        raise ValueError("The timeout keyword parameter can be either None or a number.")

    def wait_one(signal: Signal, common: Signal) -> None:
        try:
            signal.wait()
            common.turn_on()
        except CleanUp:
            pass

    common = Signal().turn_off()
    helpers = [add(wait_one, signal, common) for signal in signals]
    for helper in helpers:
        helper.hidden = True
    try:
        common.wait(timeout)
    finally:
        for helper in helpers:
            helper.interrupt(CleanUp())

    return [signal for signal in signals if signal.is_on]
