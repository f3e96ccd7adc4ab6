"""
Process-body IR: the assembler that turns a restated ITsim process body into instruction words for the engine.

ITsim bodies are arbitrary Python functions blocking inside nested calls (reference ``itsim/machine/
process_management/thread.py:41-57``), which a GPU kernel cannot run; each supported body is written
once with this assembler as an explicit state machine over the reference's own primitives.  Opcodes and operand
encodings mirror ``itsim_b200/csrc/itsb_ir.h`` one to one.
"""
from typing import Sequence, Dict, List, Optional, Tuple, Union

import numpy as np

# opcodes (itsb_ir.h)
(OP_NOP, OP_EXIT, OP_JMP, OP_LI, OP_MOV, OP_ADDI, OP_JEQ, OP_JNE, OP_JEQI, OP_JNEI, OP_DRAWI, OP_DRAW_ADDI,
 OP_ADVANCE, OP_ADVANCE_C, OP_OPEN, OP_CLOSE, OP_SEND, OP_RECV, OP_WAIT_AWAKE, OP_JASLEEP, OP_NODE_SLEEP,
 OP_NODE_WAKE, OP_SPAWN, OP_THROW, OP_LIST_CLEAR, OP_LIST_PUSH, OP_LIST_TAKE, OP_STAT, OP_RECV_FILTER,
 OP_BIND, OP_SCLOSE, OP_SENDL, OP_RECVT, OP_ADVANCE_X, OP_THREAD_EXIT, OP_CLONE, OP_FORK, OP_JOIN, OP_PWAIT, OP_TKILL,
 OP_PKILL, OP_MARK, OP_SIG_WAIT, OP_SIG_ON, OP_GSET, OP_NOW_SUB, OP_ADVANCE_X2, OP_TLOAD, OP_TSTORE, OP_ADD, OP_SUB,
 OP_JLTU, OP_XNOW, OP_XSET, OP_XELAPSE, OP_JXLE0, OP_RECVT_X, OP_SETADDR, OP_GETADDR, OP_FORWARD, OP_GETPORT,
 OP_RECV_LOG, OP_OPEN_AT, OP_SEND_LOG, OP_FORWARD_LOG, OP_CLOSE_LOG, OP_STAT_ADD, OP_STAT_FIRST,
 OP_RECV_FILTER_LOG) = range(69)

PROG_HIDDEN_BALK, PROG_HIDDEN_WAIT_ONE = 0xF1, 0xF2

# operand sources
R0, R1, R2, R3 = 0, 1, 2, 3
PKT_SRC_IP, PKT_SRC_PORT, PKT_SIZE, PKT_TAG, PKT_F0, PKT_F1, PKT_DST_IP = 0x10, 0x11, 0x12, 0x13, 0x14, 0x15, 0x16
ENT_IP, ENT_PORT = 0x18, 0x19
NODE_ADDR, NODE_INDEX, NODE_HOST, NODE_EPOCH, NODE_BCAST, SELF, EXC = 0x20, 0x21, 0x22, 0x23, 0x24, 0x25, 0x26
SELF_THREAD, SELF_TASK, PID = 0x27, 0x28, 0x29
ZERO = 0x30
G0, G1, G2, G3, G4, G5, G6, G7 = range(0x31, 0x39)
NONE = 0xFF

# OP_SEND destination modes
DST_BCAST, DST_PKT_SRC, DST_ENT, DST_SELF, DST_REGS = 0, 1, 2, 3, 4

EXC_INTERRUPT_CLASS = 0x80
EXC_HIDDEN = 0x40
EXC_ITSIM_TIMEOUT, EXC_SOCKET_CLOSED = 0x02, 0x03          # itsim.types.Timeout; ValueError("Socket is closed")
EXC_THREAD_KILLED, EXC_GS_TIMEOUT = 0x81, 0x82             # ThreadKilled(Interrupt); greensim.Timeout(Interrupt)
EXC_CLEANUP, EXC_CANCEL_BALK = 0xC1, 0xC2
NO_TIMEOUT = 0xFFFF

Target = Union[str, int]


class Asm:
    """Two-pass assembler with symbolic labels; one code array holds every program of a model."""

    def __init__(self) -> None:
        # pc 0..1 = CLOSE; EXIT: where a process lands when an Interrupt-class exception finds no handler
        self._ins: List[Tuple[int, int, int, int, object]] = [(OP_CLOSE, 0, 0, 0, 0), (OP_EXIT, 0, 0, 0, 0)]
        self._labels: Dict[str, int] = {}
        self.entries: List[Tuple[str, int]] = []   # (label, program id)
        self._entry_index: Dict[str, int] = {}
        self._scope = ""
        self._default_handler: Optional[str] = None   # inside a thread program: where an unhandled exception goes
        self.uses_machine = False
        self.log_connections = False   # socket calls assemble to their *_LOG forms (models that keep connection records)

    # -- structure ---------------------------------------------------------------------------------------
    def program(self, name: str, prog_id: int, tags: Sequence[int] = ()) -> int:
        """Starts a new program; returns its entry index.  ``tags`` are the body's greensim tags (``@tagged``,
        ``@malware``: itsim/__init__.py:9-11,52-57); processes also inherit the tags of their creator."""
        self._scope = name + "."
        self._labels[name] = len(self._ins)
        self._entry_index[name] = len(self.entries)
        mask = 0
        for t in tags:
            mask |= 1 << int(t)
        self.entries.append((name, prog_id | (mask << 16)))
        return self._entry_index[name]

    def entry(self, name: str) -> int:
        return self._entry_index[name]

    # -- shipped-itsim thread bodies -------------------------------------------------------------------------
    def thread_program(self, name: str, prog_id: int, tags: Sequence[int] = ()) -> int:
        """Starts a body run through ``Thread.run_in`` (thread.py:41-57): ``try: advance(delay); body finally:
        exit_f``.  Blocking instructions of the body default to the handler that runs ``exit_f`` and re-raises."""
        self.uses_machine = True
        idx = self.program(name, prog_id, tags)
        self._default_handler = "__exit_raise"
        self._emit(OP_ADVANCE_X, imm=("hi16", self._t("__exit_raise"), 0))
        return idx

    def end_thread_program(self) -> None:
        self._emit(OP_THREAD_EXIT, 0)
        self.label("__exit_raise")
        self._emit(OP_THREAD_EXIT, 1)
        self._default_handler = None

    def _h(self, handler: Optional[Target]) -> object:
        if handler is None and self._default_handler is not None:
            return self._t(self._default_handler)
        return self._t(handler)

    def reraise(self) -> None:
        """``raise`` inside an except/finally block of a thread body: run exit_f, then propagate."""
        self.jmp("__exit_raise")

    def add_hidden_programs(self) -> None:
        """greensim's own helper processes: the Queue.join timeout helper and select()'s
        per-signal helper.  Appended once to every model that uses the shipped-itsim primitives."""
        if "__balk" in self._entry_index:
            return
        self._default_handler = None
        self.program("__balk", PROG_HIDDEN_BALK)           # try: advance(timeout); proc.interrupt(Timeout())
        self._emit(OP_ADVANCE_X, imm=("hi16", self._t("end"), 0))
        self.throw(R0, EXC_GS_TIMEOUT)                     # except CancelBalk: pass
        self.label("end")
        self.exit()
        self.program("__wait_one", PROG_HIDDEN_WAIT_ONE)   # try: signal.wait(); common.turn_on()
        self._emit(OP_SIG_WAIT, R0, imm=("hi16", self._t("end"), NO_TIMEOUT))
        self._emit(OP_SIG_ON, R1)                          # except CleanUp: pass
        self.label("end")
        self.exit()

    def label(self, name: str) -> None:
        key = self._scope + name
        if key in self._labels:
            raise ValueError(f"duplicate label {key}")
        self._labels[key] = len(self._ins)

    def _emit(self, op: int, a: int = 0, b: int = 0, c: int = 0, imm: object = 0) -> None:
        self._ins.append((op, a, b, c, imm))

    def _t(self, target: Optional[Target]) -> object:
        if target is None:
            return 0
        if isinstance(target, int):
            return target
        if target.startswith("/"):
            return ("label", target[1:])        # absolute: "/program.label"
        return ("label", self._scope + target)

    # -- instructions --------------------------------------------------------------------------------------
    def exit(self) -> None: self._emit(OP_EXIT)
    def jmp(self, target: Target) -> None: self._emit(OP_JMP, imm=self._t(target))
    def li(self, rd: int, value: int) -> None: self._emit(OP_LI, rd, imm=value & 0xFFFFFFFF)
    def mov(self, rd: int, src: int) -> None: self._emit(OP_MOV, rd, src)
    def addi(self, rd: int, src: int, value: int) -> None: self._emit(OP_ADDI, rd, src, imm=value & 0xFFFFFFFF)
    def jeq(self, a: int, b: int, target: Target) -> None: self._emit(OP_JEQ, a, b, imm=self._t(target))
    def jne(self, a: int, b: int, target: Target) -> None: self._emit(OP_JNE, a, b, imm=self._t(target))
    def jeqi(self, a: int, value: int, target: Target) -> None: self._emit(OP_JEQI, a, imm=("lo16", self._t(target), value))
    def jnei(self, a: int, value: int, target: Target) -> None: self._emit(OP_JNEI, a, imm=("lo16", self._t(target), value))
    def drawi(self, rd: int, dist: int) -> None: self._emit(OP_DRAWI, rd, imm=dist)
    def draw_addi(self, rd: int, src: int, dist: int) -> None: self._emit(OP_DRAW_ADDI, rd, src, imm=dist)

    def advance(self, dist: int, handler: Optional[Target] = None) -> None:
        self._emit(OP_ADVANCE, imm=("hi16", self._h(handler), dist))

    def advance_const(self, const_index: int, handler: Optional[Target] = None) -> None:
        self._emit(OP_ADVANCE_C, imm=("hi16", self._h(handler), const_index))

    # shipped-itsim primitives
    def bind(self, port: int = 0, detached: bool = False) -> None:
        """``detached``: the socket outlives the binding process (``run_networking_daemon`` binds at set-up and
        hands the socket to a chain of threads, node.py:340-348)."""
        self._emit(OP_BIND, 1 if detached else 0, imm=port)
    def sclose(self) -> None: self._emit(OP_SCLOSE)

    def sendl(self, dst_mode: int, size_src: int, tag: int, port: int = 0, f0: int = NONE, f1: int = NONE,
              handler: Optional[Target] = None) -> None:
        self._emit(OP_SENDL, dst_mode, size_src, tag, (port & 0xFFFF) | (f0 << 16) | (f1 << 24))

    def recvt(self, timeout_const: int = NO_TIMEOUT, handler: Optional[Target] = None) -> None:
        self._emit(OP_RECVT, imm=("hi16", self._h(handler), timeout_const))

    def clone(self, program: str, delay_const: int, rd: int = NONE) -> None:
        self._emit(OP_CLONE, rd, imm=("hi16", delay_const, ("entry", program)))

    def fork(self, program: str, delay_const: int, rd: int = NONE) -> None:
        self._emit(OP_FORK, rd, imm=("hi16", delay_const, ("entry", program)))

    def join(self, thread_src: int, timeout_const: int = NO_TIMEOUT, handler: Optional[Target] = None) -> None:
        self._emit(OP_JOIN, thread_src, imm=("hi16", self._h(handler), timeout_const))

    def pwait(self, task_src: int, timeout_const: int = NO_TIMEOUT, handler: Optional[Target] = None) -> None:
        self._emit(OP_PWAIT, task_src, imm=("hi16", self._h(handler), timeout_const))

    def tkill(self, thread_src: int) -> None: self._emit(OP_TKILL, thread_src)
    def pkill(self, task_src: int) -> None: self._emit(OP_PKILL, task_src)
    def mark(self, code: int) -> None: self._emit(OP_MARK, NONE, imm=code)
    def mark_value(self, src: int) -> None: self._emit(OP_MARK, src)
    def getport(self, rd: int) -> None: self._emit(OP_GETPORT, rd)
    def gset(self, index: int, src: int) -> None: self._emit(OP_GSET, src, imm=index)
    def now_sub(self, const_index: int) -> None: self._emit(OP_NOW_SUB, imm=const_index)
    def advance_x2(self, handler: Optional[Target] = None) -> None:
        self._emit(OP_ADVANCE_X2, imm=("hi16", self._h(handler), 0))

    def tload(self, rd: int, index_src: int, base: int = 0) -> None: self._emit(OP_TLOAD, rd, index_src, imm=base)
    def tstore(self, value_src: int, index_src: int, base: int = 0) -> None: self._emit(OP_TSTORE, value_src, index_src, imm=base)
    def add(self, rd: int, a: int, b: int) -> None: self._emit(OP_ADD, rd, a, b)
    def sub(self, rd: int, a: int, b: int) -> None: self._emit(OP_SUB, rd, a, b)
    def jltu(self, a: int, b: int, target: Target) -> None: self._emit(OP_JLTU, a, b, imm=self._t(target))
    def xnow(self) -> None: self._emit(OP_XNOW)
    def xset(self, const_index: int) -> None: self._emit(OP_XSET, imm=const_index)
    def xelapse(self) -> None: self._emit(OP_XELAPSE)
    def jxle0(self, target: Target) -> None: self._emit(OP_JXLE0, imm=self._t(target))

    def recvt_x(self, handler: Optional[Target] = None) -> None:
        self._emit(OP_RECVT_X, imm=("hi16", self._h(handler), 0))

    def forward(self, from_node: object, from_port: int) -> None:
        """Relay the packet just received from another attachment of the machine (node object resolved at
        compile time) to (r2, r3)."""
        self._emit(OP_FORWARD_LOG if self.log_connections else OP_FORWARD, imm=("node_port", from_node, from_port))

    def open_at(self, node: object, port: int) -> None:
        """``open_socket((address of another attachment of this machine, port))``: a socket FORWARD sends from."""
        self._emit(OP_OPEN_AT, imm=("node_port", node, port))

    def setaddr(self, src: int) -> None: self._emit(OP_SETADDR, src)
    def getaddr(self, rd: int) -> None: self._emit(OP_GETADDR, rd)

    def open(self, port: int = 0) -> None: self._emit(OP_OPEN, imm=port)
    def close(self) -> None: self._emit(OP_CLOSE_LOG if self.log_connections else OP_CLOSE)

    def send(self, dst_mode: int, size_src: int, tag: int, port: int = 0, f0: int = NONE, f1: int = NONE) -> None:
        self._emit(OP_SEND_LOG if self.log_connections else OP_SEND, dst_mode, size_src, tag,
                   (port & 0xFFFF) | (f0 << 16) | (f1 << 24))

    def recv(self, handler: Optional[Target] = None) -> None:
        self._emit(OP_RECV_LOG if self.log_connections else OP_RECV, imm=("hi16", self._t(handler), 0))

    def recv_filter(self, rd: int, filter_index: int, asleep: Target, handler: Optional[Target] = None) -> None:
        """Fused ``while True: with ws.awake(): packet = socket.recv()`` loop head that also consumes the packets
        the program's own dispatch would ignore (see ``Simulator.packet_filter``)."""
        self._emit(OP_RECV_FILTER_LOG if self.log_connections else OP_RECV_FILTER, rd, filter_index,
                   imm=("hi16", self._t(handler), self._t(asleep)))

    def wait_awake(self, rd: int = NONE, handler: Optional[Target] = None) -> None:
        self._emit(OP_WAIT_AWAKE, rd, imm=("hi16", self._t(handler), 0))

    def jasleep(self, src: int, target: Target) -> None: self._emit(OP_JASLEEP, src, imm=self._t(target))
    def node_sleep(self) -> None: self._emit(OP_NODE_SLEEP)
    def node_wake(self) -> None: self._emit(OP_NODE_WAKE)
    def spawn(self, program: str, rd: int = NONE) -> None: self._emit(OP_SPAWN, rd, imm=("entry", program))
    def throw(self, handle_src: int, exc: int) -> None: self._emit(OP_THROW, handle_src, imm=exc)
    def list_clear(self) -> None: self._emit(OP_LIST_CLEAR)
    def list_push(self) -> None: self._emit(OP_LIST_PUSH)
    def list_take(self, key_src: int, not_found: Target) -> None: self._emit(OP_LIST_TAKE, key_src, imm=self._t(not_found))
    def stat(self, counter: int) -> None: self._emit(OP_STAT, imm=counter)
    def stat_add(self, counter: int, src: int) -> None: self._emit(OP_STAT_ADD, src, imm=counter)

    def stat_first(self, counter: int) -> None:
        """counter = now() in microseconds, the first time this runs in a replica (never 0: a replica where it
        never ran keeps 0)."""
        self._emit(OP_STAT_FIRST, imm=counter)

    # -- output ----------------------------------------------------------------------------------------------
    def _resolve(self, v: object) -> int:
        if isinstance(v, int):
            return v
        if isinstance(v, tuple):
            if v[0] == "label":
                if v[1] not in self._labels:
                    raise KeyError(f"undefined label {v[1]}")
                return self._labels[v[1]]
            if v[0] == "entry":
                return self._entry_index[v[1]]
            if v[0] == "hi16":
                return (self._resolve(v[2]) & 0xFFFF) | (self._resolve(v[1]) << 16)
            if v[0] == "lo16":
                return (self._resolve(v[1]) & 0xFFFF) | ((self._resolve(v[2]) & 0xFFFF) << 16)
            if v[0] == "node_port":
                return (v[1].index & 0xFFFF) | ((v[2] & 0xFFFF) << 16)
        raise TypeError(v)

    def assemble(self) -> Tuple[np.ndarray, np.ndarray]:
        """Returns (code uint64[n], entries (pc, prog) uint32[n_entries, 2])."""
        code = np.zeros(len(self._ins), dtype=np.uint64)
        for pc, (op, a, b, c, imm) in enumerate(self._ins):
            if pc == 0 and self.log_connections:
                op = OP_CLOSE_LOG      # the landing pad of an unhandled Interrupt closes the socket too
            word = (op & 0xFF) | ((a & 0xFF) << 8) | ((b & 0xFF) << 16) | ((c & 0xFF) << 24) \
                | ((self._resolve(imm) & 0xFFFFFFFF) << 32)
            code[pc] = word
        for pc in range(len(code)):       # operand c of a fused loop head: 1 = its loop is stable (see _stable_loop)
            if int(code[pc]) & 0xFF == OP_RECV_FILTER and self._stable_loop(code, pc):
                code[pc] = np.uint64(int(code[pc]) | (1 << 24))
        entries = np.zeros((len(self.entries), 2), dtype=np.uint32)
        for i, (name, prog) in enumerate(self.entries):
            entries[i] = (self._labels[name], prog)
        return code, entries

    # what a program may do with a packet its loop head hands over and still have a "stable" loop: nothing that blocks,
    # ends the process, or touches the socket other than sending -- so that the packets queued behind are certain to be
    # popped by the same loop head in the same activation
    _STABLE_OPS = frozenset((OP_NOP, OP_LI, OP_MOV, OP_ADDI, OP_DRAWI, OP_DRAW_ADDI, OP_SEND, OP_FORWARD, OP_LIST_CLEAR,
                             OP_LIST_PUSH, OP_STAT, OP_STAT_ADD, OP_STAT_FIRST))

    @classmethod
    def _stable_loop(cls, code: np.ndarray, head: int) -> bool:
        """Does every path from the instruction after the RECV_FILTER at ``head`` come back to ``head`` through
        non-blocking instructions only?  (The engine then knows that a packet the filter is bound to reject can be
        counted instead of queued while the workstation sleeps: itsb_core.cuh, ND_ASLEEP.)"""
        seen, todo = set(), [head + 1]
        while todo:
            pc = todo.pop()
            if pc == head or pc in seen:
                continue
            if pc >= len(code):
                return False
            seen.add(pc)
            word = int(code[pc])
            op, imm = word & 0xFF, word >> 32
            if op in cls._STABLE_OPS:
                todo.append(pc + 1)
            elif op == OP_JMP:
                todo.append(imm & 0xFFFF)
            elif op in (OP_JEQ, OP_JNE, OP_JASLEEP, OP_LIST_TAKE):
                todo += [pc + 1, imm & 0xFFFF]
            elif op in (OP_JEQI, OP_JNEI):
                todo += [pc + 1, imm & 0xFFFF]
            else:
                return False
        return True
