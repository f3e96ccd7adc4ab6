"""
Model front-end: compiles ITsim process bodies written as plain Python functions into engine programs.

The reference runs a body as a coroutine (``Thread.run_in``, itsim/machine/process_management/thread.py:41-57); the
engine runs instruction words (``itsb_ir.h``).  This module closes the gap for bodies in the style of the reference's
own models and integration tests (tests/integ/test_packet_transfer.py, test_broadcast.py, test_thread_lifecycle.py,
test_process_lifecycle.py): it reads the function's source (``inspect`` + ``ast``) and lowers it with *partial
evaluation* --

* everything that does not depend on the simulation (module constants, ``self`` attributes, ``1 * S``, loops over
  Python lists, object construction) is evaluated here, at compile time, in the function's own globals / closure;
* the simulation primitives become instructions: ``with context.node.bind(...) as socket`` (BIND / SCLOSE),
  ``socket.send`` (SENDL), ``socket.recv(timeout)`` (RECVT), ``advance`` (ADVANCE*), ``context.thread.clone[_in]``
  (CLONE), ``context.process.fork_exec[_in]`` (FORK), ``.join`` / ``.wait`` / ``.kill``, ``try / except Timeout /
  finally`` (handler chains), ``while`` / ``if`` on packet fields and handles;
* statements whose only effect is on *host* objects (``ledger.add("client")``, ``cemetary.append(name)``,
  ``pytest.fail(...)``, ``assert now() > 5.0``) become MARK records; after a run, :meth:`Frontend.replay` re-executes
  them on the host in trace order with ``now()`` bound to the record's time, so the model's Python-side state ends up
  as the reference leaves it.

Up to four simulation values (handles, counters) can be live in a body -- the engine's process has four registers;
a body that needs more raises :class:`CompileError`, as does any construct outside the subset.
"""
import ast
import builtins
import inspect
import textwrap
from ipaddress import IPv4Address
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

from . import ir
from .random import Dist

MARK_EFFECT_BASE = 0x1000            # MARK codes at and above this are front-end effects (index = code - base)

_EXC_BY_NAME = {"Timeout": ir.EXC_ITSIM_TIMEOUT, "ThreadKilled": ir.EXC_THREAD_KILLED,
                "ValueError": ir.EXC_SOCKET_CLOSED}


class CompileError(Exception):
    """The body uses something outside the supported subset; the message names the line."""


def intrinsic(name: str) -> Callable:
    def deco(f: Callable) -> Callable:
        f.__itsb_intrinsic__ = name
        return f
    return deco


def _intrinsic_name(obj: Any) -> Optional[str]:
    """``now`` / ``advance`` are recognised whether they come from this package's compatibility modules or from the
    reference's ``itsim.simulator`` (which re-exports greensim's)."""
    tag = getattr(obj, "__itsb_intrinsic__", None)
    if tag:
        return tag
    mod = getattr(obj, "__module__", "") or ""
    if inspect.isfunction(obj) and obj.__name__ in ("now", "advance") and mod.split(".")[0] in ("greensim", "itsim"):
        return obj.__name__
    return None


# ---- compile-time values ---------------------------------------------------------------------------------------------
class Const:
    __slots__ = ("v",)

    def __init__(self, v: Any) -> None:
        self.v = v


class Rt:
    """A value that exists only in the running replica: an operand source of the IR plus what it denotes."""
    __slots__ = ("src", "kind")

    def __init__(self, src: int, kind: str = "int") -> None:
        self.src, self.kind = src, kind


class PerProcess:
    """Marks an argument of ``run_proc`` as a per-process value: it travels in a register instead of being folded into
    the program, so that one compiled program serves every node it is started on (``per_process(peer_address)``)."""

    def __init__(self, value: Any) -> None:
        self.value = value


def per_process(value: Any) -> PerProcess:
    return PerProcess(value)


class GlobalRef:
    """What a shared Python dict holds, at compile time, for an entry stored from a running body: model variable g[i]."""

    def __init__(self, slot: int, kind: str) -> None:
        self.slot, self.kind = slot, kind


class _Sym:
    def __init__(self, what: str, **kw: Any) -> None:
        self.what = what
        self.__dict__.update(kw)


class Effect:
    """One host-side statement lifted out of a body."""

    def __init__(self, code: Any, env: Dict[str, Any], text: str, value_kind: Optional[str], lineno: int) -> None:
        self.code, self.env, self.text, self.value_kind, self.lineno = code, env, text, value_kind, lineno

    def run(self, t: float, value: Optional[int]) -> None:
        env = dict(self.env)
        env["__itsb_now__"] = lambda: t
        if self.value_kind is not None:
            env["__itsb_value__"] = IPv4Address(value) if self.value_kind == "address" else value
        exec(self.code, env)


class Frontend:
    """Per-``Simulator`` compiler state: compiled programs, host effects, payload tags, shared variables."""

    def __init__(self, sim: Any) -> None:
        self.sim = sim
        self.asm: ir.Asm = sim.asm
        self.program_ids: Dict[str, int] = {}        # function name -> program id (what the trace reports); optional
        self.effects: List[Effect] = []
        self.content_tags: Dict[Any, int] = {}
        self.payload_fields: Dict[Any, int] = {}     # payload key -> packet field 0 / 1
        self._programs: Dict[tuple, Tuple[str, int]] = {}
        self._queue: List[Tuple[str, Callable, tuple, int]] = []
        self._gslots: Dict[Tuple[int, Any], int] = {}
        self._keepalive: List[Any] = []
        self._next_prog = 1
        self._gtemps = 0
        self._mutated: Dict[int, str] = {}           # id of a host object -> the statement of a body that modifies it
        self._steering: Dict[int, Tuple[str, str]] = {}   # id of a host object read at compile time -> (where, text)
        self.single_interface_programs: set = set()     # programs that use the "every non-loopback interface" idiom

    # -- public ----------------------------------------------------------------------------------------------------
    def program(self, f: Callable, args: Sequence[Any]) -> Tuple[str, List[int]]:
        """(name of the program for body ``f`` started with ``args``, initial register values).  Arguments are
        compile-time values, folded into the program, except those wrapped in :func:`per_process`."""
        regs, lowered = [], []
        for a in args:
            if isinstance(a, PerProcess):
                v = a.value
                kind = "address" if isinstance(v, (IPv4Address, str)) else "int"
                regs.append(int(IPv4Address(v)) if kind == "address" else int(v) & 0xFFFFFFFF)
                lowered.append(Rt(0, kind))
            else:
                lowered.append(Const(a))
        name, _ = self._request(f, tuple(lowered))
        self._drain()
        return name, regs

    def tag_of(self, content: Any) -> int:
        if content not in self.content_tags:
            self.content_tags[content] = len(self.content_tags) + 1
            if self.content_tags[content] > 255:
                raise CompileError("more than 255 distinct payload contents")
        return self.content_tags[content]

    def payload_field(self, key: Any, body: Any, node: ast.AST) -> int:
        """Which of a packet's two 32-bit fields carries payload key ``key`` (assigned on first use, model-wide)."""
        if key not in self.payload_fields:
            if len(self.payload_fields) >= 2:
                raise body.err(node, f"payload keys {sorted(map(str, list(self.payload_fields) + [key]))} are not carried "
                                     f"by the engine's packets (a packet carries 'content' and two 32-bit fields)")
            self.payload_fields[key] = len(self.payload_fields)
        return self.payload_fields[key]

    def replay(self, trace: Any) -> int:
        """Re-executes the host-side statements of the bodies in the order (and at the simulated times) the replica
        executed them.  Returns how many ran.  Exceptions they raise (``pytest.fail``, a failed ``assert``) propagate."""
        from . import _lib as L
        n, pending = 0, None
        for rec in trace:
            if int(rec["kind"]) != L.TRACE_MARK:
                continue
            aux = int(rec["aux"])
            if pending is not None:
                pending.run(float(rec["t"]), aux)
                pending, n = None, n + 1
                continue
            if aux < MARK_EFFECT_BASE or aux - MARK_EFFECT_BASE >= len(self.effects):
                continue
            eff = self.effects[aux - MARK_EFFECT_BASE]
            if eff.value_kind is not None:
                pending = eff
            else:
                eff.run(float(rec["t"]), None)
                n += 1
        return n

    def effect_text(self, mark_code: int) -> Optional[str]:
        i = mark_code - MARK_EFFECT_BASE
        return self.effects[i].text if 0 <= i < len(self.effects) else None

    # -- internals -------------------------------------------------------------------------------------------------
    def _request(self, f: Callable, args: Tuple[Any, ...]) -> Tuple[str, List[int]]:
        """Registers (f, compile-time arguments) and returns (program name, positions of the run-time arguments)."""
        rt_positions = [i for i, a in enumerate(args) if isinstance(a, Rt)]
        key = (getattr(f, "__func__", f), id(getattr(f, "__self__", None)),
               tuple(("rt", a.kind) if isinstance(a, Rt) else (("sym", a.what) if isinstance(a, _Sym) else ("c", id(a.v)))
                     for a in args))
        if key not in self._programs:
            base = f.__name__
            name, k = base, 1
            taken = {n for n, _ in self._programs.values()}
            while name in taken:
                k += 1
                name = f"{base}#{k}"
            prog_id = self.program_ids.get(base)
            if prog_id is None:
                prog_id = self._next_prog
                self.program_ids[base] = prog_id
            self._next_prog = max(self._next_prog, prog_id + 1)
            self._programs[key] = (name, prog_id)
            self._keepalive.append((f, args))
            self._queue.append((name, f, args, prog_id))
        return self._programs[key][0], rt_positions

    def _drain(self) -> None:
        while self._queue:
            name, f, args, prog_id = self._queue.pop(0)
            _Body(self, name, f, args, prog_id).compile()
        self.check_host_state()

    def check_host_state(self) -> None:
        """Host objects the bodies modify (``ledger.add(...)``, at replay) must not also steer the compilation
        (``if len(ledger) > 3:`` would be decided once, here, on the object as it is now): that would be a silent
        difference from the reference, where the body sees the live object."""
        for obj_id, (where, text) in self._steering.items():
            if obj_id in self._mutated:
                raise CompileError(f"{where}: `{text}` reads a host object that {self._mutated[obj_id]} modifies while the "
                                   f"simulation runs; host state changed by a body cannot steer the simulation")

    def gtemp(self, i: int) -> int:
        self._gtemps = max(self._gtemps, i + 1)
        if len(self._gslots) + self._gtemps > 8:
            raise CompileError("more than 8 simulation values shared between bodies")
        return 7 - i

    def gslot(self, container: Any, key: Any) -> int:
        k = (id(container), key)
        if k not in self._gslots:
            if len(self._gslots) + self._gtemps >= 8:
                raise CompileError("more than 8 simulation values shared between bodies")
            self._gslots[k] = len(self._gslots)
            self._keepalive.append(container)
        return self._gslots[k]

    def add_effect(self, eff: Effect) -> int:
        self.effects.append(eff)
        return MARK_EFFECT_BASE + len(self.effects) - 1


# ---- one body ----------------------------------------------------------------------------------------------------------
class _Scope:
    """Names of one (possibly inlined) function activation."""

    def __init__(self, f: Callable, parent_env: Optional[Dict[str, Any]] = None) -> None:
        func = getattr(f, "__func__", f)
        self.globals = func.__globals__
        self.names: Dict[str, Any] = {}
        if func.__closure__:
            for n, cell in zip(func.__code__.co_freevars, func.__closure__):
                try:
                    self.names[n] = Const(cell.cell_contents)
                except ValueError:
                    pass
        self.end_label: Optional[str] = None
        self.fin_depth = 0

    def lookup(self, name: str) -> Any:
        if name in self.names:
            return self.names[name]
        if name in self.globals:
            return Const(self.globals[name])
        if hasattr(builtins, name):
            return Const(getattr(builtins, name))
        raise NameError(name)


class _Body:

    def __init__(self, fe: Frontend, name: str, f: Callable, args: Tuple[Any, ...], prog_id: int) -> None:
        self.fe, self.a, self.sim = fe, fe.asm, fe.sim
        self.name, self.f, self.args, self.prog_id = name, f, args, prog_id
        self.n_labels = 0
        self.handlers: List[str] = []            # innermost last; "" = the thread's own exit handler
        self.loops: List[Tuple[str, str, int]] = []   # (continue label, break label, handler depth)
        self.finalizers: List[Tuple[str, Any, int]] = []   # what `return` / `break` / `continue` must run on their way out
        self.packet_gen = 0
        self.activation = 0                       # bumped by every blocking call: packet fields do not survive one
        self.low = 0                              # registers [0, low) hold parameters (and locals in send mode)
        self.high = 4                             # registers [high, 4) hold locals (default mode)
        self.scratch_used: set = set()
        self.send_mode = False

    # -- helpers -----------------------------------------------------------------------------------------------------
    def err(self, node: ast.AST, msg: str) -> CompileError:
        return CompileError(f"{self.f.__name__}, line {getattr(node, 'lineno', '?')}: {msg}")

    def label(self, hint: str) -> str:
        self.n_labels += 1
        return f"_{hint}{self.n_labels}"

    def handler(self) -> Optional[str]:
        return self.handlers[-1] if self.handlers else None

    def alloc_local(self, node: ast.AST) -> int:
        if self.send_mode:
            r = self.low
            self.low += 1
        else:
            self.high -= 1
            r = self.high
        if self.low > self.high or (self.send_mode and self.low > 2):
            raise self.err(node, "more simulation values live in this body than the engine's process has registers")
        return r

    def scratch(self, node: ast.AST, avoid: Sequence[int] = ()) -> int:
        order = (1, 0) if self.send_mode else (0, 1, 2, 3)
        for r in order:
            if self.low <= r < self.high and r not in avoid:
                if self.send_mode and r >= 2:
                    continue
                self.scratch_used.add(r)
                return r
        raise self.err(node, "no free register for a temporary value")

    def check_registers(self) -> None:
        for r in self.scratch_used:
            if not (self.low <= r < self.high):
                raise CompileError(f"{self.f.__name__}: too many simulation values live at once (register r{r} is both a "
                                   f"temporary and a variable)")

    def src_of(self, node: ast.AST, v: Any, avoid: Sequence[int] = ()) -> int:
        """Operand source holding ``v`` (loads constants into a temporary register)."""
        if isinstance(v, Rt):
            return v.src
        if isinstance(v, Const):
            val = v.v
            if isinstance(val, GlobalRef):
                return ir.G0 + val.slot
            val = self.as_u32(node, val)
            if val == 0:
                return ir.ZERO
            r = self.scratch(node, avoid)
            self.a.li(r, val)
            return r
        raise self.err(node, "expected a number or a simulation value")

    def as_u32(self, node: ast.AST, val: Any) -> int:
        if isinstance(val, bool):
            return int(val)
        if isinstance(val, int):
            return val & 0xFFFFFFFF
        if isinstance(val, IPv4Address):
            return int(val)
        if isinstance(val, str):
            try:
                return int(IPv4Address(val))
            except ValueError:
                return self.fe.tag_of(val)
        raise self.err(node, f"cannot represent {val!r} as a 32-bit simulation value")

    # -- compile -----------------------------------------------------------------------------------------------------
    def compile(self) -> None:
        func = getattr(self.f, "__func__", self.f)
        try:
            src = textwrap.dedent(inspect.getsource(func))
        except (OSError, TypeError) as exc:
            raise CompileError(f"{self.f!r}: source not available ({exc})")
        tree = ast.parse(src)
        ast.increment_lineno(tree, func.__code__.co_firstlineno - 1)      # errors name the line in the user's file
        fdef = tree.body[0]
        if not isinstance(fdef, (ast.FunctionDef,)):
            raise CompileError(f"{self.f!r}: not a plain function")
        for n in ast.walk(fdef):
            if isinstance(n, (ast.Yield, ast.YieldFrom, ast.Await, ast.Lambda, ast.FunctionDef)) and n is not fdef:
                raise self.err(n, f"{type(n).__name__} is not supported in a body")
        self.send_mode = any(isinstance(n, ast.Attribute) and n.attr == "send" for n in ast.walk(fdef)) and \
            self._has_explicit_destination(fdef)
        scope = _Scope(self.f)
        self.bind_parameters(fdef, scope, self.f, list(self.args), top=True)
        tags = getattr(func, "_tag_set", None) or getattr(func, "__itsb_tags__", ())
        self.a.thread_program(self.name, self.prog_id, tags=[int(getattr(t, "value", t)) for t in tags])
        scope.end_label = self.label("return")
        self.block(fdef.body, scope)
        self.a.label(scope.end_label)
        self.a.end_thread_program()
        self.check_registers()

    @staticmethod
    def _has_explicit_destination(fdef: ast.AST) -> bool:
        for n in ast.walk(fdef):
            if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr == "send" and n.args:
                d = n.args[0]
                if not (isinstance(d, ast.Attribute) and d.attr == "source"):
                    return True
        return False

    def bind_parameters(self, fdef: ast.FunctionDef, scope: _Scope, f: Callable, args: List[Any], top: bool) -> None:
        params = [p.arg for p in fdef.args.args]
        if getattr(f, "__self__", None) is not None:
            scope.names[params.pop(0)] = Const(f.__self__)
        if not params:
            raise CompileError(f"{f.__name__}: a body takes the context as its first parameter")
        scope.names[params.pop(0)] = _Sym("context")
        if len(args) > len(params):
            raise CompileError(f"{f.__name__}: too many arguments")
        defaults = fdef.args.defaults
        for i, p in enumerate(params):
            if i < len(args):
                v = args[i]
                if isinstance(v, Rt) and top:
                    v = Rt(self.low, v.kind)          # run-time arguments arrive in r0.. in order
                    self.low += 1
                scope.names[p] = v
            else:
                j = i - (len(params) - len(defaults))
                if j < 0:
                    raise CompileError(f"{f.__name__}: missing argument {p}")
                scope.names[p] = Const(ast.literal_eval(defaults[j]))

    # -- statements ----------------------------------------------------------------------------------------------------
    def block(self, stmts: Sequence[ast.stmt], scope: _Scope) -> None:
        for s in stmts:
            self.stmt(s, scope)

    def stmt(self, s: ast.stmt, scope: _Scope) -> None:
        if isinstance(s, ast.Pass):
            return
        if isinstance(s, ast.Expr):
            if isinstance(s.value, ast.Constant):      # docstring
                return
            return self.expr_stmt(s, scope)
        if isinstance(s, (ast.Assign, ast.AnnAssign)):
            return self.assign(s, scope)
        if isinstance(s, ast.With):
            return self.with_(s, scope)
        if isinstance(s, ast.Try):
            return self.try_(s, scope)
        if isinstance(s, ast.While):
            return self.while_(s, scope)
        if isinstance(s, ast.For):
            return self.for_(s, scope)
        if isinstance(s, ast.If):
            return self.if_(s, scope)
        if isinstance(s, ast.Assert):
            return self.assert_(s, scope)
        if isinstance(s, ast.Return):
            if s.value is not None and not (isinstance(s.value, ast.Constant) and s.value.value is None):
                raise self.err(s, "a body's return value is not used by the simulation")
            self.unwind(scope.fin_depth, s)
            self.a.jmp(scope.end_label)
            return
        if isinstance(s, ast.Break):
            if not self.loops:
                raise self.err(s, "break outside a loop")
            self.unwind(self.loops[-1][2], s)
            self.a.jmp(self.loops[-1][1])
            return
        if isinstance(s, ast.Continue):
            if not self.loops:
                raise self.err(s, "continue outside a loop")
            self.unwind(self.loops[-1][2], s)
            self.a.jmp(self.loops[-1][0])
            return
        if isinstance(s, ast.Raise):
            if s.exc is None:
                return self.propagate()
            raise self.err(s, "raising a new exception inside a body is not supported (bare `raise` is)")
        raise self.err(s, f"statement {type(s).__name__} is not supported in a body")

    def unwind(self, depth: int, node: ast.AST) -> None:
        """Runs, innermost first, what leaving the enclosing with / try-finally blocks down to ``depth`` entails."""
        for i in range(len(self.finalizers) - 1, depth - 1, -1):
            kind, payload, h_depth = self.finalizers[i]
            if kind == "close":
                self.a.sclose()
            else:
                body, scope = payload
                saved_f, saved_h = self.finalizers, self.handlers
                self.finalizers, self.handlers = saved_f[:i], saved_h[:h_depth]
                self.block(body, scope)
                self.finalizers, self.handlers = saved_f, saved_h

    def propagate(self) -> None:
        """The pending exception goes to the next enclosing handler (or out of the thread: exit_f, re-raise)."""
        outer = self.handlers[-1] if self.handlers else None
        if outer is None:
            self.a.reraise()
        else:
            self.a.jmp(outer)

    # with context.node.bind(...) as socket:
    def with_(self, s: ast.With, scope: _Scope) -> None:
        if len(s.items) != 1:
            raise self.err(s, "one context manager per with statement")
        item = s.items[0]
        self._bind_in_with = True
        try:
            v = self.eval(item.context_expr, scope)
        finally:
            self._bind_in_with = False
        if not (isinstance(v, _Sym) and v.what == "socket"):
            raise self.err(s, "only `with context.node.bind(...) as socket` is supported")
        if item.optional_vars is not None:
            if not isinstance(item.optional_vars, ast.Name):
                raise self.err(s, "with ... as <name>")
            scope.names[item.optional_vars.id] = v
        h = self.label("with_exc")
        self.finalizers.append(("close", None, len(self.handlers)))
        self.handlers.append(h)
        self.block(s.body, scope)
        self.finalizers.pop()
        self.handlers.pop()
        self.a.sclose()
        end = self.label("with_end")
        self.a.jmp(end)
        self.a.label(h)
        self.a.sclose()
        self.propagate()
        self.a.label(end)

    def try_(self, s: ast.Try, scope: _Scope) -> None:
        end = self.label("try_end")
        h_fin = self.label("finally_exc") if s.finalbody else None
        if h_fin:
            self.finalizers.append(("finally", (s.finalbody, scope), len(self.handlers)))
            self.handlers.append(h_fin)
        h_exc = self.label("except") if s.handlers else None
        if h_exc:
            self.handlers.append(h_exc)
        self.block(s.body, scope)
        if h_exc:
            self.handlers.pop()
        self.block(s.orelse, scope)
        after = self.label("try_ok")
        self.a.jmp(after)
        if h_exc:
            self.a.label(h_exc)
            for hd in s.handlers:
                nxt = self.label("next_except")
                codes = self.exception_codes(hd, scope)
                if codes is not None:
                    if len(codes) != 1:
                        raise self.err(hd, "one exception class per except clause")
                    self.a.jnei(ir.EXC, codes[0], nxt)
                if hd.name:
                    scope.names[hd.name] = _Sym("exception")
                self.block(hd.body, scope)
                self.a.jmp(after)
                self.a.label(nxt)
            self.propagate()
        self.a.label(after)
        if h_fin:
            self.finalizers.pop()
            self.handlers.pop()
            self.block(s.finalbody, scope)
            self.a.jmp(end)
            self.a.label(h_fin)
            self.block(s.finalbody, scope)
            self.propagate()
        self.a.label(end)

    def exception_codes(self, hd: ast.ExceptHandler, scope: _Scope) -> Optional[List[int]]:
        if hd.type is None:
            return None
        v = self.eval(hd.type, scope)
        classes = v.v if isinstance(v, Const) else None
        if classes is Exception or classes is BaseException:
            return None
        if not isinstance(classes, tuple):
            classes = (classes,)
        out = []
        for c in classes:
            code = _EXC_BY_NAME.get(getattr(c, "__name__", ""))
            if code is None:
                raise self.err(hd, f"exception class {c!r} has no engine counterpart")
            out.append(code)
        return out

    def while_(self, s: ast.While, scope: _Scope) -> None:
        if s.orelse:
            raise self.err(s, "while ... else")
        top, end = self.label("loop"), self.label("loop_end")
        self.a.label(top)
        cond = self.eval(s.test, scope)
        if isinstance(cond, Const):
            self.note_steering(s.test, scope)
            if not cond.v:
                return
        else:
            self.branch_unless(s.test, cond, end)
        self.loops.append((top, end, len(self.finalizers)))
        self.block(s.body, scope)
        self.loops.pop()
        self.a.jmp(top)
        self.a.label(end)

    def for_(self, s: ast.For, scope: _Scope) -> None:
        it = self.eval(s.iter, scope)
        if not isinstance(it, Const):
            raise self.err(s, "a for loop in a body iterates over compile-time values (it is unrolled)")
        self.note_steering(s.iter, scope)
        if s.orelse:
            raise self.err(s, "for ... else")
        end = self.label("for_end")
        for item in list(it.v):
            nxt = self.label("for_next")
            self.bind_target(s.target, item if isinstance(item, _Sym) else Const(item), scope, s)
            self.loops.append((nxt, end, len(self.finalizers)))
            self.block(s.body, scope)
            self.loops.pop()
            self.a.label(nxt)
        self.a.label(end)

    def if_(self, s: ast.If, scope: _Scope) -> None:
        cond = self.eval(s.test, scope)
        if isinstance(cond, Const):
            self.note_steering(s.test, scope)
            return self.block(s.body if cond.v else s.orelse, scope)
        other, end = self.label("else"), self.label("endif")
        self.branch_unless(s.test, cond, other)
        self.block(s.body, scope)
        self.a.jmp(end)
        self.a.label(other)
        self.block(s.orelse, scope)
        self.a.label(end)

    def branch_unless(self, node: ast.AST, cond: Any, target: str) -> None:
        """Jumps to ``target`` when the run-time condition is false."""
        if not (isinstance(cond, _Sym) and cond.what == "compare"):
            raise self.err(node, "a run-time condition must be a comparison (==, !=, <, >=, is, is not)")
        op, l, r = cond.op, cond.left, cond.right
        if op in ("lt", "ge", "gt", "le"):
            if op in ("gt", "le"):
                l, r, op = r, l, {"gt": "lt", "le": "ge"}[op]
            ls = self.src_of(node, l)
            rs = self.src_of(node, r, avoid=[ls])
            if op == "lt":          # unless l < r: skip over a jump
                skip = self.label("lt")
                self.a.jltu(ls, rs, skip)
                self.a.jmp(target)
                self.a.label(skip)
            else:
                self.a.jltu(ls, rs, target)
            return
        if isinstance(l, Const) and isinstance(r, Rt):
            l, r = r, l
        if isinstance(r, Const) and not isinstance(r.v, GlobalRef):
            val = self.as_u32(node, r.v)
            if val < 0x10000:
                (self.a.jnei if op == "eq" else self.a.jeqi)(self.src_of(node, l), val, target)
                return
        ls = self.src_of(node, l)
        rs = self.src_of(node, r, avoid=[ls])
        (self.a.jne if op == "eq" else self.a.jeq)(ls, rs, target)

    def assert_(self, s: ast.Assert, scope: _Scope) -> None:
        if self.is_host_only(s.test, scope):
            return self.host_effect(s, scope)
        cond = self.eval(s.test, scope)
        if isinstance(cond, Const):
            if not cond.v:
                raise self.err(s, "assertion is false at compile time")
            return
        ok = self.label("assert_ok")
        bad = self.label("assert_bad")
        self.branch_unless(s.test, cond, bad)
        self.a.jmp(ok)
        self.a.label(bad)
        text = ast.unparse(s.test)
        code = compile(ast.parse(f"raise AssertionError({text!r})"), f"<{self.f.__name__}:{s.lineno}>", "exec")
        self.a.mark(self.fe.add_effect(Effect(code, {}, f"assert {text}", None, s.lineno)))
        self.a.label(ok)

    # -- host-side statements --------------------------------------------------------------------------------------------
    def is_host_only(self, node: ast.AST, scope: _Scope) -> bool:
        """True when the expression reads nothing of the running replica except the clock."""
        for n in ast.walk(node):
            if isinstance(n, ast.Name):
                try:
                    v = scope.lookup(n.id)
                except NameError:
                    return False
                if not isinstance(v, Const) or isinstance(v.v, GlobalRef):
                    return False
            if isinstance(n, ast.Subscript):
                # a shared dict may hold simulation values
                try:
                    v = self.eval(n, scope)
                except Exception:
                    return False
                if not isinstance(v, Const) or isinstance(v.v, GlobalRef):
                    return False
        return True

    def snapshot_env(self, scope: _Scope) -> Dict[str, Any]:
        env = dict(scope.globals)
        for k, v in scope.names.items():
            if isinstance(v, Const):
                env[k] = v.v
        return env

    @staticmethod
    def _is_mutable(obj: Any) -> bool:
        if isinstance(obj, (list, set, dict, bytearray)):
            return True
        return hasattr(obj, "__dict__") and not isinstance(obj, (type, type(ast), type(len))) and \
            not inspect.isfunction(obj) and not inspect.ismethod(obj)

    def note_mutation(self, s: ast.stmt, scope: _Scope) -> None:
        """The receiver of a method call statement, and the mutable arguments of a plain call, are what the statement
        may modify."""
        call = s.value if isinstance(s, ast.Expr) and isinstance(s.value, ast.Call) else None
        if call is None:
            return
        targets = [call.func.value] if isinstance(call.func, ast.Attribute) else list(call.args)
        for t in targets:
            try:
                v = self.eval(t, scope)
            except CompileError:
                continue
            if isinstance(v, Const) and self._is_mutable(v.v):
                self.fe._mutated.setdefault(id(v.v), f"`{ast.unparse(s)}` ({self.f.__name__}, line {s.lineno})")
                self.fe._keepalive.append(v.v)

    def note_steering(self, e: ast.AST, scope: _Scope) -> None:
        """Every host object a compile-time decision looks at."""
        for n in ast.walk(e):
            if isinstance(n, (ast.Name, ast.Attribute, ast.Subscript)):
                try:
                    v = self.eval(n, scope)
                except Exception:
                    continue
                if isinstance(v, Const) and self._is_mutable(v.v):
                    self.fe._steering.setdefault(id(v.v), (f"{self.f.__name__}, line {getattr(e, 'lineno', '?')}",
                                                           ast.unparse(e)))
                    self.fe._keepalive.append(v.v)

    def host_effect(self, s: ast.stmt, scope: _Scope, value: Optional[Rt] = None, value_node: Optional[ast.AST] = None) -> None:
        """Lifts statement ``s`` to the host: a MARK now, the statement itself at replay."""
        self.note_mutation(s, scope)

        class Rewrite(ast.NodeTransformer):
            def visit(self_inner, n: ast.AST) -> ast.AST:      # noqa: N805
                if value_node is not None and n is value_node:
                    return ast.copy_location(ast.Name("__itsb_value__", ast.Load()), n)
                return super().visit(n)

            def visit_Call(self_inner, n: ast.Call) -> ast.AST:      # noqa: N805
                self_inner.generic_visit(n)
                if isinstance(n.func, ast.Name):
                    try:
                        tgt = scope.lookup(n.func.id)
                    except NameError:
                        return n
                    if isinstance(tgt, Const) and _intrinsic_name(tgt.v) == "now":
                        n.func = ast.copy_location(ast.Name("__itsb_now__", ast.Load()), n.func)
                return n

        tree = ast.Module(body=[Rewrite().visit(_copy(s, value_node))], type_ignores=[])
        ast.fix_missing_locations(tree)
        code = compile(tree, f"<{self.f.__name__}:{s.lineno}>", "exec")
        eff = Effect(code, self.snapshot_env(scope), ast.unparse(s), value.kind if value is not None else None, s.lineno)
        self.a.mark(self.fe.add_effect(eff))
        if value is not None:
            self.a.mark_value(value.src)

    def expr_stmt(self, s: ast.Expr, scope: _Scope) -> None:
        call = s.value
        if not isinstance(call, ast.Call):
            raise self.err(s, "expression statement without effect")
        if self.is_host_only(call, scope):
            fv = self.eval(call.func, scope)
            if isinstance(fv, Const) and self.inlinable(fv.v):
                return self.inline_call(s, fv.v, [self.eval(x, scope) for x in call.args], scope)
            if isinstance(fv, Const) and _intrinsic_name(fv.v) == "advance":
                return self.advance(call, scope)
            return self.host_effect(s, scope)
        fv = self.eval(call.func, scope)
        if isinstance(fv, Const) and _intrinsic_name(fv.v) == "advance":
            return self.advance(call, scope)
        if isinstance(fv, Const) and (inspect.isfunction(fv.v) or inspect.ismethod(fv.v)):
            args = [self.eval(x, scope) for x in call.args]
            if self.inlinable(fv.v, args):
                return self.inline_call(s, fv.v, args, scope)
        if isinstance(fv, _Sym) and fv.what == "method":
            self.sim_call(call, fv, scope, want_result=False)
            return
        if isinstance(fv, Const):
            # a host call taking exactly one simulation value: container.add(packet.source.hostname_as_address())
            rt_args = [(x, self.eval(x, scope)) for x in call.args]
            rts = [(x, v) for x, v in rt_args if isinstance(v, Rt)]
            if len(rts) == 1 and all(isinstance(v, (Const, Rt)) for _, v in rt_args):
                return self.host_effect(s, scope, value=rts[0][1], value_node=rts[0][0])
        raise self.err(s, "this call mixes host objects and simulation values in a way the front-end cannot split")

    def inlinable(self, f: Any, args: Sequence[Any] = ()) -> bool:
        """A Python function is compiled in place when it takes the context first (``self._live(context, child)``), or
        when it is handed the context, a socket or a packet (``daemon.trigger(context, packet, socket)``)."""
        if not (inspect.isfunction(f) or inspect.ismethod(f)):
            return False
        func = getattr(f, "__func__", f)
        if _intrinsic_name(func) is not None:
            return False
        mod = (getattr(func, "__module__", "") or "").split(".")[0]
        if getattr(func, "__itsb_inline__", False):
            return True
        if mod in ("pytest", "_pytest", "builtins", "greensim", "itsim", "itsim_b200"):
            return False
        if any(isinstance(a, _Sym) and a.what in ("context", "socket", "packet") for a in args):
            return True
        try:
            params = list(inspect.signature(func).parameters)
        except (TypeError, ValueError):
            return False
        if inspect.ismethod(f):
            params = params[1:]
        return bool(params) and params[0] in ("context", "ctx")

    def inline_call(self, s: ast.stmt, f: Callable, args: List[Any], scope: _Scope) -> None:
        """A body calling a helper that takes the context (``self._live(context, child)``): compiled in place."""
        func = getattr(f, "__func__", f)
        tree = ast.parse(textwrap.dedent(inspect.getsource(func)))
        ast.increment_lineno(tree, func.__code__.co_firstlineno - 1)
        fdef = tree.body[0]
        inner = _Scope(f)
        params = [p.arg for p in fdef.args.args]
        if getattr(f, "__self__", None) is not None:
            inner.names[params.pop(0)] = Const(f.__self__)
        if len(args) > len(params):
            raise self.err(s, "too many arguments")
        for p, v in zip(params, args):
            inner.names[p] = v
        inner.end_label = self.label("inline_return")
        inner.fin_depth = len(self.finalizers)
        self.block(fdef.body, inner)
        self.a.label(inner.end_label)

    # -- assignments -----------------------------------------------------------------------------------------------------
    def assign(self, s: Any, scope: _Scope) -> None:
        if isinstance(s, ast.AnnAssign):
            if s.value is None:
                return
            targets = [s.target]
        else:
            targets = s.targets
        if len(targets) != 1:
            raise self.err(s, "chained assignment")
        tgt = targets[0]
        def is_action(fv: Any) -> bool:      # a call that emits instructions (not a read of a packet field)
            return isinstance(fv, _Sym) and fv.what == "method" and isinstance(fv.owner, _Sym) and \
                fv.owner.what in ("socket", "thread_self", "task_self", "node")

        if isinstance(tgt, ast.Name) and isinstance(s.value, ast.Call):
            fv = self.eval(s.value.func, scope)
            if is_action(fv):
                existing = scope.names.get(tgt.id)
                rd = existing.src if isinstance(existing, Rt) and existing.src < 4 else None
                v = self.sim_call(s.value, fv, scope, want_result=True, name=tgt.id, rd=rd, node=s)
                scope.names[tgt.id] = v
                return
        if isinstance(tgt, ast.Subscript) and isinstance(s.value, ast.Call):
            fv = self.eval(s.value.func, scope)
            if is_action(fv):
                cont, key = self.eval(tgt.value, scope), self.eval(tgt.slice, scope)
                if not (isinstance(cont, Const) and isinstance(key, Const)):
                    raise self.err(s, "container and key must be compile-time values")
                r = self.scratch(s)
                v = self.sim_call(s.value, fv, scope, want_result=True, rd=r, node=s)
                slot = self.fe.gslot(cont.v, key.v)
                self.a.gset(slot, r)
                cont.v[key.v] = GlobalRef(slot, v.kind)
                return
        v = self.eval(s.value, scope)
        self.bind_target(tgt, v, scope, s)

    def bind_target(self, tgt: ast.AST, v: Any, scope: _Scope, node: ast.AST) -> None:
        if isinstance(tgt, ast.Name):
            if isinstance(v, Rt) and v.src >= 4 and v.src not in (ir.SELF_TASK, ir.SELF_THREAD, ir.PID) \
                    and not (ir.G0 <= v.src <= ir.G7):
                # packet fields change with the next recv: copy
                existing = scope.names.get(tgt.id)
                r = existing.src if isinstance(existing, Rt) and existing.src < 4 else self.alloc_local(node)
                self.a.mov(r, v.src)
                v = Rt(r, v.kind)
            scope.names[tgt.id] = v
            return
        if isinstance(tgt, ast.Tuple) and isinstance(v, Const):
            items = list(v.v)
            if len(items) != len(tgt.elts):
                raise self.err(node, "unpacking length mismatch")
            for t, item in zip(tgt.elts, items):
                self.bind_target(t, Const(item), scope, node)
            return
        if isinstance(tgt, ast.Subscript):
            cont, key = self.eval(tgt.value, scope), self.eval(tgt.slice, scope)
            if isinstance(cont, Const) and isinstance(key, Const):
                if isinstance(v, Const):
                    cont.v[key.v] = v.v
                    return
                if isinstance(v, Rt):
                    slot = self.fe.gslot(cont.v, key.v)
                    self.a.gset(slot, v.src)
                    cont.v[key.v] = GlobalRef(slot, v.kind)
                    return
        if isinstance(tgt, ast.Attribute):
            obj = self.eval(tgt.value, scope)
            if isinstance(obj, Const) and isinstance(v, Const):
                setattr(obj.v, tgt.attr, v.v)
                return
        raise self.err(node, "unsupported assignment target")

    # -- simulation calls --------------------------------------------------------------------------------------------------
    def timeout_const(self, call: ast.Call, scope: _Scope, pos: int = 0) -> int:
        node = call.args[pos] if len(call.args) > pos else None
        for kw in call.keywords:
            if kw.arg == "timeout":
                node = kw.value
        if node is None:
            return ir.NO_TIMEOUT
        v = self.eval(node, scope)
        if not isinstance(v, Const):
            raise self.err(call, "a timeout must be a compile-time number")
        if v.v is None:
            return ir.NO_TIMEOUT
        return self.sim.const(float(v.v))

    def advance(self, call: ast.Call, scope: _Scope) -> None:
        if len(call.args) != 1:
            raise self.err(call, "advance(delay)")
        self.blocked()
        arg = call.args[0]
        # advance(next(generator))
        if isinstance(arg, ast.Call) and isinstance(arg.func, ast.Name) and arg.func.id == "next" and len(arg.args) == 1:
            g = self.eval(arg.args[0], scope)
            if isinstance(g, Const) and isinstance(g.v, Dist):
                return self.a.advance(self.sim.dist(g.v), handler=self.handler())
            raise self.err(call, "advance(next(g)): g must be one of this package's generator descriptors")
        # advance(T - now())
        if isinstance(arg, ast.BinOp) and isinstance(arg.op, ast.Sub) and isinstance(arg.right, ast.Call):
            fv = self.eval(arg.right.func, scope)
            if isinstance(fv, Const) and _intrinsic_name(fv.v) == "now":
                left = self.eval(arg.left, scope)
                if not isinstance(left, Const):
                    raise self.err(call, "advance(T - now()): T must be a compile-time number")
                self.a.now_sub(self.sim.const(float(left.v)))
                return self.a.advance_x2(handler=self.handler())
        v = self.eval(arg, scope)
        if isinstance(v, Const) and isinstance(v.v, (int, float)):
            return self.a.advance_const(self.sim.const(float(v.v)), handler=self.handler())
        raise self.err(call, "advance(): the delay must be a compile-time number, next(generator) or T - now()")

    def sim_call(self, call: ast.Call, m: _Sym, scope: _Scope, want_result: bool, name: Optional[str] = None,
                 rd: Optional[int] = None, node: Optional[ast.AST] = None) -> Any:
        owner, attr = m.owner, m.attr
        node = node or call
        what = owner.what if isinstance(owner, _Sym) else owner.kind
        if what == "node" and attr == "bind":
            return self.bind(call, scope)
        if what == "socket":
            if attr == "send":
                return self.send(call, scope)
            if attr == "recv":
                self.a.recvt(self.timeout_const(call, scope), handler=self.handler())
                self.packet_gen += 1
                self.blocked()
                return _Sym("packet", gen=self.packet_gen, act=self.activation)
            if attr == "close":
                return self.a.sclose()
            raise self.err(call, f"Socket.{attr} is not supported")
        if what in ("thread_self", "thread") and attr in ("clone", "clone_in"):
            if what != "thread_self":
                raise self.err(call, "clone is called on context.thread")
            return self.spawn(call, scope, attr == "clone_in", self.a.clone, "thread", rd, node)
        if what in ("task_self", "task") and attr in ("fork_exec", "fork_exec_in"):
            if what != "task_self":
                raise self.err(call, "fork_exec is called on context.process")
            return self.spawn(call, scope, attr == "fork_exec_in", self.a.fork, "task", rd, node)
        src = {"thread_self": ir.SELF_THREAD, "task_self": ir.SELF_TASK}.get(what)
        if src is None:
            src = owner.src
        kind = "thread" if what.startswith("thread") else "task"
        if attr == "join" and kind == "thread":
            self.blocked()
            return self.a.join(src, self.timeout_const(call, scope), handler=self.handler())
        if attr == "wait" and kind == "task":
            self.blocked()
            return self.a.pwait(src, self.timeout_const(call, scope), handler=self.handler())
        if attr == "kill":
            return (self.a.tkill if kind == "thread" else self.a.pkill)(src)
        raise self.err(call, f"{what}.{attr} is not supported")

    def send(self, call: ast.Call, scope: _Scope) -> None:
        args = list(call.args)
        kw = {k.arg: k.value for k in call.keywords}
        dest_node = args[0] if args else kw.get("dr")
        size_node = args[1] if len(args) > 1 else kw.get("size")
        payload_node = args[2] if len(args) > 2 else kw.get("payload")
        if dest_node is None or size_node is None:
            raise self.err(call, "socket.send(destination, size[, payload])")
        tag, fields = 0, {}
        if payload_node is not None:
            if isinstance(payload_node, ast.Dict):
                items = {}
                for k, v in zip(payload_node.keys, payload_node.values):
                    kv = self.eval(k, scope) if k is not None else None
                    if not isinstance(kv, Const):
                        raise self.err(call, "payload keys are compile-time values")
                    items[kv.v] = self.eval(v, scope)
            else:
                p = self.eval(payload_node, scope)
                if not isinstance(p, Const) or not (p.v is None or isinstance(p.v, dict)):
                    raise self.err(call, "a payload is a dict")
                items = {k: Const(v) for k, v in (p.v or {}).items()}
            for key, v in items.items():
                if key == "content":
                    if not isinstance(v, Const):
                        raise self.err(call, "payload['content'] is a compile-time value")
                    tag = self.fe.tag_of(v.v)
                    continue
                if not isinstance(v, (Const, Rt)) or (isinstance(v, Const) and not isinstance(v.v, (int, str, IPv4Address))):
                    raise self.err(call, f"payload keys {[key]} are not carried by the engine's packets (a packet "
                                         f"carries 'content' and two 32-bit fields)")
                fields[self.fe.payload_field(key, self, call)] = v
        fsrc = dict(fields)
        f0 = self.field_src(call, fsrc.get(0))
        f1 = self.field_src(call, fsrc.get(1), avoid=[f0])
        fkw = dict(f0=f0, f1=f1)
        if isinstance(dest_node, ast.Tuple) and len(dest_node.elts) == 2:
            addr, port = self.eval(dest_node.elts[0], scope), self.eval(dest_node.elts[1], scope)
            size = self.eval(size_node, scope)
            if isinstance(addr, _Sym) and addr.what == "broadcast_address":
                if not isinstance(port, Const):
                    raise self.err(call, "the port of a broadcast is a compile-time value")
                self.fe.single_interface_programs.add(self.name)
                return self.a.sendl(ir.DST_BCAST, self.src_of(size_node, size, avoid=[f0, f1]), tag, port=int(port.v), **fkw)
            if isinstance(addr, (Rt, Const)) and isinstance(port, (Rt, Const)):
                return self.send_regs(call, addr, port, size, tag, fkw)
        dest = self.eval(dest_node, scope)
        size = self.eval(size_node, scope)
        if isinstance(dest, _Sym) and dest.what == "packet_source":
            self.need_current(dest_node, dest)
            return self.a.sendl(ir.DST_PKT_SRC, self.src_of(size_node, size, avoid=[f0, f1]), tag, **fkw)
        if isinstance(dest, Const) and isinstance(dest.v, tuple) and len(dest.v) == 2:
            return self.send_regs(call, Const(dest.v[0]), Const(dest.v[1]), size, tag, fkw)
        raise self.err(call, "destination must be packet.source or an (address, port) pair")

    def is_live(self, r: int) -> bool:
        """Register r holds a parameter or a variable of this body."""
        return r < self.low or r >= self.high

    def send_regs(self, node: ast.AST, addr: Any, port: Any, size: Any, tag: int, fkw: Dict[str, int]) -> None:
        """SENDL with the destination in (r2, r3).  Variables living in r2 / r3, and operands that would be overwritten
        while the pair is loaded, are parked in model variables reserved for this (nothing runs in between) and put
        back after the send."""
        def park(i: int, src: int) -> int:
            slot = self.fe.gtemp(2 + i)
            self.a.gset(slot, src)
            return ir.G0 + slot

        want = {ir.R2: addr, ir.R3: port}
        srcs: Dict[int, Any] = {}
        for reg, v in want.items():
            srcs[reg] = v.src if isinstance(v, Rt) else None
        size_src = size.src if isinstance(size, Rt) else None
        fsrc = {k: v for k, v in fkw.items()}
        saved: Dict[int, int] = {}
        for i, reg in enumerate((ir.R2, ir.R3)):
            if srcs[reg] == reg:
                continue                                  # already in place
            if self.is_live(reg) or any(s_ == reg for s_ in (srcs[ir.R2], srcs[ir.R3], size_src, *fsrc.values())):
                g = park(i, reg)
                if self.is_live(reg):
                    saved[reg] = g
                # operands that were read from this register are read from the parked copy instead
                for other in (ir.R2, ir.R3):
                    if srcs[other] == reg:
                        srcs[other] = g
                if size_src == reg:
                    size_src = g
                for k in fsrc:
                    if fsrc[k] == reg:
                        fsrc[k] = g
            else:
                self.scratch_used.add(reg)
        for reg, v in want.items():
            if isinstance(v, Rt):
                if srcs[reg] != reg:
                    self.a.mov(reg, srcs[reg])
            else:
                self.a.li(reg, self.as_u32(node, v.v) if reg == ir.R2 else int(v.v))
        if size_src is None:
            val = self.as_u32(node, size.v)
            if val == 0:
                size_src = ir.ZERO
            else:
                free = [r for r in (0, 1) if not self.is_live(r) and r not in fsrc.values()]
                if free:
                    self.scratch_used.add(free[0])
                    self.a.li(free[0], val)
                    size_src = free[0]
                else:                                     # every register is taken: through a parked constant
                    r = ir.R3
                    g_port = park(2, r)
                    self.a.li(r, val)
                    slot = self.fe.gtemp(5)
                    self.a.gset(slot, r)
                    self.a.mov(r, g_port)
                    size_src = ir.G0 + slot
        self.a.sendl(ir.DST_REGS, size_src, tag, **fsrc)
        for reg, g in saved.items():
            self.a.mov(reg, g)

    def field_src(self, node: ast.AST, v: Any, avoid: Sequence[int] = ()) -> int:
        """Operand source of a payload field (NONE: the packet does not carry it -> 0)."""
        if v is None:
            return ir.NONE
        if isinstance(v, Rt):
            return v.src
        val = self.as_u32(node, v.v)
        if val == 0:
            return ir.ZERO
        # a constant field is parked in a model variable reserved for this (g7 / g6): the registers stay free for the
        # destination and the size; nothing can run between the store and the send
        slot = self.fe.gtemp(0 if not any(a == ir.G7 for a in avoid) else 1)
        r = self.scratch(node)
        self.a.li(r, val)
        self.a.gset(slot, r)
        return ir.G0 + slot

    def need_current(self, node: ast.AST, sym: _Sym) -> None:
        if sym.gen != self.packet_gen:
            raise self.err(node, "this packet is no longer the last one received (the engine keeps one per process)")
        if getattr(sym, "act", self.activation) != self.activation:
            raise self.err(node, "packet fields do not survive a blocking call (advance, recv, join, wait): copy what is "
                                 "needed into variables first (size = packet.byte_size; port = packet.source.port; ...)")

    def blocked(self) -> None:
        self.activation += 1

    def spawn(self, call: ast.Call, scope: _Scope, timed: bool, emit: Callable, kind: str, rd: Optional[int],
              node: ast.AST) -> Any:
        args = list(call.args)
        delay = 0.0
        if timed:
            d = self.eval(args.pop(0), scope)
            if not isinstance(d, Const):
                raise self.err(call, "the start delay must be a compile-time number")
            delay = float(d.v)
        if not args:
            raise self.err(call, "missing body function")
        fv = self.eval(args.pop(0), scope)
        if not (isinstance(fv, Const) and callable(fv.v)):
            raise self.err(call, "the body must be a compile-time function")
        vals = []
        for x in args:
            v = self.eval(x, scope)
            if isinstance(v, Const) and isinstance(v.v, GlobalRef):
                v = Rt(ir.G0 + v.v.slot, v.v.kind)
            if isinstance(v, _Sym):
                if v.what in ("task_self", "thread_self"):
                    v = Rt(ir.SELF_TASK if v.what == "task_self" else ir.SELF_THREAD, v.what.split("_")[0])
                elif v.what == "socket":
                    if emit != self.a.clone:
                        raise self.err(call, "a socket can be handed to a thread of the same process (clone) only")
                elif v.what != "context":
                    raise self.err(call, f"cannot pass a {v.what} to another body")
            vals.append(v)
        callee_args = tuple(Rt(0, v.kind) if isinstance(v, Rt) else v for v in vals)
        name, rt_positions = self.fe._request(fv.v, callee_args)
        # the child starts with a copy of this process's registers: its run-time parameters are r0.. in order
        for reg, pos in enumerate(rt_positions):
            v = vals[pos]
            if v.src != reg:
                if not (self.low <= reg < self.high):
                    raise self.err(call, f"argument {pos + 1} has to travel in r{reg}, which holds another live value")
                self.scratch_used.add(reg)
                self.a.mov(reg, v.src)
        if rd is None and isinstance(node, (ast.Assign, ast.AnnAssign)):
            rd = self.alloc_local(node)
        emit(name, self.sim.const(delay), rd=ir.NONE if rd is None else rd)
        return Rt(rd, kind) if rd is not None else None

    # -- expressions -------------------------------------------------------------------------------------------------------
    def eval(self, e: ast.AST, scope: _Scope) -> Any:
        if isinstance(e, ast.Constant):
            return Const(e.value)
        if isinstance(e, ast.Name):
            try:
                return scope.lookup(e.id)
            except NameError:
                raise self.err(e, f"name {e.id!r} is not defined")
        if isinstance(e, ast.Attribute):
            return self.attribute(e, self.eval(e.value, scope), scope)
        if isinstance(e, ast.Subscript):
            base, key = self.eval(e.value, scope), self.eval(e.slice, scope)
            if isinstance(base, _Sym) and base.what == "payload":
                self.need_current(e, base)
                if isinstance(key, Const) and key.v == "content":
                    return Rt(ir.PKT_TAG, "content")
                if isinstance(key, Const):
                    return Rt(ir.PKT_F0 if self.fe.payload_field(key.v, self, e) == 0 else ir.PKT_F1, "int")
                raise self.err(e, "payload keys are compile-time values")
            if isinstance(base, Const) and isinstance(key, Const):
                return Const(base.v[key.v])
            raise self.err(e, "subscript on a simulation value")
        if isinstance(e, ast.Call):
            return self.call(e, scope)
        if isinstance(e, ast.Compare):
            return self.compare(e, scope)
        if isinstance(e, ast.UnaryOp) and isinstance(e.op, ast.Not):
            v = self.eval(e.operand, scope)
            if isinstance(v, Const):
                return Const(not v.v)
            if isinstance(v, _Sym) and v.what == "compare":
                inv = {"eq": "ne", "ne": "eq", "lt": "ge", "ge": "lt", "gt": "le", "le": "gt"}
                return _Sym("compare", op=inv[v.op], left=v.left, right=v.right)
            raise self.err(e, "not <simulation value>")
        if isinstance(e, (ast.Tuple, ast.List, ast.Set)):
            items = [self.eval(x, scope) for x in e.elts]
            if all(isinstance(i, Const) for i in items):
                seq = [i.v for i in items]
                return Const(tuple(seq) if isinstance(e, ast.Tuple) else (set(seq) if isinstance(e, ast.Set) else seq))
            raise self.err(e, "a tuple / list of simulation values")
        if isinstance(e, ast.Dict):
            ks = [self.eval(k, scope) for k in e.keys]
            vs = [self.eval(v, scope) for v in e.values]
            if all(isinstance(i, Const) for i in ks + vs):
                return Const({k.v: v.v for k, v in zip(ks, vs)})
            raise self.err(e, "a dict of simulation values")
        # anything else: compile-time only
        return self.static(e, scope)

    def static(self, e: ast.AST, scope: _Scope) -> Const:
        if not self.is_host_only(e, scope):
            raise self.err(e, f"expression `{ast.unparse(e)}` mixes simulation values into Python arithmetic")
        code = compile(ast.Expression(body=e), f"<{self.f.__name__}:{getattr(e, 'lineno', 0)}>", "eval")
        try:
            return Const(eval(code, self.snapshot_env(scope)))
        except Exception as exc:
            raise self.err(e, f"compile-time evaluation of `{ast.unparse(e)}` failed: {exc!r}")

    def attribute(self, e: ast.Attribute, base: Any, scope: _Scope) -> Any:
        attr = e.attr
        if isinstance(base, Const):
            if isinstance(base.v, GlobalRef):
                base = Rt(ir.G0 + base.v.slot, base.v.kind)
            else:
                try:
                    return Const(getattr(base.v, attr))
                except AttributeError as exc:
                    raise self.err(e, str(exc))
        if isinstance(base, Rt):
            if base.kind in ("thread", "task") and attr in ("join", "wait", "kill"):
                return _Sym("method", owner=base, attr=attr)
            raise self.err(e, f"attribute {attr!r} of a simulation value")
        what = base.what
        if what == "context":
            if attr == "node":
                return _Sym("node")
            if attr == "process":
                return _Sym("task_self")
            if attr == "thread":
                return _Sym("thread_self")
        elif what == "node":
            if attr in ("bind", "interfaces"):
                return _Sym("method", owner=base, attr=attr)
        elif what == "interface":
            if attr == "address":
                return _Sym("interface_address")
            if attr == "cidr":
                return _Sym("interface_cidr")
        elif what == "interface_address":
            if attr == "is_loopback":
                return Const(False)
        elif what == "interface_cidr":
            if attr == "broadcast_address":
                return _Sym("broadcast_address")
        elif what == "task_self":
            if attr == "pid":
                return Rt(ir.PID, "int")
            if attr == "node":
                return _Sym("node")
            if attr in ("fork_exec", "fork_exec_in", "kill", "wait"):
                return _Sym("method", owner=base, attr=attr)
        elif what == "thread_self":
            if attr == "process":
                return _Sym("task_self")
            if attr in ("clone", "clone_in", "kill", "join"):
                return _Sym("method", owner=base, attr=attr)
        elif what == "socket":
            if attr in ("send", "recv", "close"):
                return _Sym("method", owner=base, attr=attr)
        elif what == "packet":
            self.need_current(e, base)
            if attr == "byte_size":
                return Rt(ir.PKT_SIZE, "int")
            if attr == "source":
                return _Sym("packet_source", gen=base.gen, act=base.act)
            if attr == "dest":
                return _Sym("packet_dest", gen=base.gen, act=base.act)
            if attr == "payload":
                return _Sym("payload", gen=base.gen, act=base.act)
        elif what == "packet_source":
            self.need_current(e, base)
            if attr == "port":
                return Rt(ir.PKT_SRC_PORT, "int")
            if attr == "hostname_as_address":
                return _Sym("method", owner=base, attr=attr)
            if attr == "hostname":
                return Rt(ir.PKT_SRC_IP, "address")
        elif what == "packet_dest":
            self.need_current(e, base)
            if attr in ("hostname_as_address",):
                return _Sym("method", owner=base, attr=attr)
            if attr == "hostname":
                return Rt(ir.PKT_DST_IP, "address")
        elif what == "payload":
            if attr == "get":
                return _Sym("method", owner=base, attr="get")
        raise self.err(e, f"{what}.{attr} is not supported")

    def call(self, e: ast.Call, scope: _Scope) -> Any:
        fv = self.eval(e.func, scope)
        if isinstance(fv, _Sym) and fv.what == "method":
            owner = fv.owner
            what = owner.what if isinstance(owner, _Sym) else owner.kind
            if what == "node" and fv.attr == "bind":
                return self.bind(e, scope)
            if what == "node" and fv.attr == "interfaces":
                # `for interface in context.node.interfaces(): if not interface.address.is_loopback: ...` in a program
                # shared by many nodes: the loopback is skipped here, the one other interface stands for itself
                return Const([_Sym("interface")])
            if what in ("packet_source", "packet_dest") and fv.attr == "hostname_as_address":
                return Rt(ir.PKT_SRC_IP if what == "packet_source" else ir.PKT_DST_IP, "address")
            if what == "payload" and fv.attr == "get":
                key = self.eval(e.args[0], scope)
                if isinstance(key, Const) and key.v == "content":
                    return Rt(ir.PKT_TAG, "content")
                if isinstance(key, Const):
                    return Rt(ir.PKT_F0 if self.fe.payload_field(key.v, self, e) == 0 else ir.PKT_F1, "int")
                raise self.err(e, "payload keys are compile-time values")
            if what == "socket" and fv.attr == "recv":
                return self.sim_call(e, fv, scope, want_result=True)
            raise self.err(e, f"{what}.{fv.attr}(...) cannot be used as a value here")
        if isinstance(fv, Const):
            if _intrinsic_name(fv.v) == "now":
                raise self.err(e, "now() can be used in host-side statements (asserts, ledger updates) and in "
                                  "advance(T - now()) only")
            return self.static(e, scope)
        raise self.err(e, "call on a simulation value")

    def bind(self, e: ast.Call, scope: _Scope) -> _Sym:
        args = [self.eval(x, scope) for x in e.args]
        kw = {k.arg: self.eval(k.value, scope) for k in e.keywords}
        port = args[1] if len(args) > 1 else kw.get("pr", Const(0))
        if not isinstance(port, Const):
            raise self.err(e, "the port must be a compile-time value")
        p = port.v
        if p is None:
            p = 0
        if isinstance(p, (tuple, list, range)):
            raise self.err(e, "port ranges are not supported")
        # outside a `with`, the socket is not closed when this thread ends: Node.run_networking_daemon binds it once and
        # hands it from thread to thread (node.py:340-348)
        self.a.bind(int(p), detached=not getattr(self, "_bind_in_with", False))
        return _Sym("socket")

    def compare(self, e: ast.Compare, scope: _Scope) -> Any:
        if len(e.ops) != 1:
            return self.static(e, scope)
        l, r = self.eval(e.left, scope), self.eval(e.comparators[0], scope)
        conv = {"task_self": Rt(ir.SELF_TASK, "task"), "thread_self": Rt(ir.SELF_THREAD, "thread")}
        l = conv.get(l.what, l) if isinstance(l, _Sym) else l
        r = conv.get(r.what, r) if isinstance(r, _Sym) else r
        if isinstance(l, Const) and isinstance(r, Const) and not isinstance(l.v, GlobalRef) \
                and not isinstance(r.v, GlobalRef):
            return self.static(e, scope)
        if isinstance(l, _Sym) or isinstance(r, _Sym):
            raise self.err(e, "comparison of objects the engine has no value for")
        op = {ast.Eq: "eq", ast.NotEq: "ne", ast.Is: "eq", ast.IsNot: "ne", ast.Lt: "lt", ast.GtE: "ge", ast.Gt: "gt",
              ast.LtE: "le"}.get(type(e.ops[0]))
        if op is None:
            raise self.err(e, "unsupported comparison operator")
        return _Sym("compare", op=op, left=l, right=r)


def _copy(s: ast.stmt, value_node: Optional[ast.AST]) -> ast.stmt:
    """A deep copy of the statement in which ``value_node`` keeps its identity (the rewrite finds it by identity)."""
    import copy
    return copy.deepcopy(s, {id(value_node): value_node} if value_node is not None else {})
