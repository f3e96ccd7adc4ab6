"""
Process-body IR: the assembler that turns a restated ITsim process body into instruction words for the engine.

ITsim bodies are arbitrary Python functions blocking inside nested calls (reference ``itsim/machine/
process_management/thread.py:41-57``), which a GPU kernel cannot run; each supported body is written
once with this assembler as an explicit state machine over the reference's own primitives.  Opcodes and operand
encodings mirror ``itsim_b200/csrc/itsb_ir.h`` one to one.
"""
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

# opcodes (itsb_ir.h)
(OP_NOP, OP_EXIT, OP_JMP, OP_LI, OP_MOV, OP_ADDI, OP_JEQ, OP_JNE, OP_JEQI, OP_JNEI, OP_DRAWI, OP_DRAW_ADDI,
 OP_ADVANCE, OP_ADVANCE_C, OP_OPEN, OP_CLOSE, OP_SEND, OP_RECV, OP_WAIT_AWAKE, OP_JASLEEP, OP_NODE_SLEEP,
 OP_NODE_WAKE, OP_SPAWN, OP_THROW, OP_LIST_CLEAR, OP_LIST_PUSH, OP_LIST_TAKE, OP_STAT, OP_RECV_FILTER) = range(29)

# operand sources
R0, R1, R2, R3 = 0, 1, 2, 3
PKT_SRC_IP, PKT_SRC_PORT, PKT_SIZE, PKT_TAG, PKT_F0, PKT_F1 = 0x10, 0x11, 0x12, 0x13, 0x14, 0x15
ENT_IP, ENT_PORT = 0x18, 0x19
NODE_ADDR, NODE_INDEX, NODE_HOST, NODE_EPOCH, NODE_BCAST, SELF, EXC = 0x20, 0x21, 0x22, 0x23, 0x24, 0x25, 0x26
ZERO = 0x30
NONE = 0xFF

# OP_SEND destination modes
DST_BCAST, DST_PKT_SRC, DST_ENT, DST_SELF, DST_REGS = 0, 1, 2, 3, 4

EXC_INTERRUPT_CLASS = 0x80

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

    # -- structure ---------------------------------------------------------------------------------------
    def program(self, name: str, prog_id: int) -> int:
        """Starts a new program; returns its entry index."""
        self._scope = name + "."
        self._labels[name] = len(self._ins)
        self._entry_index[name] = len(self.entries)
        self.entries.append((name, prog_id))
        return self._entry_index[name]

    def entry(self, name: str) -> int:
        return self._entry_index[name]

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
        self._emit(OP_ADVANCE, imm=("hi16", self._t(handler), dist))

    def advance_const(self, const_index: int, handler: Optional[Target] = None) -> None:
        self._emit(OP_ADVANCE_C, imm=("hi16", self._t(handler), const_index))

    def open(self, port: int = 0) -> None: self._emit(OP_OPEN, imm=port)
    def close(self) -> None: self._emit(OP_CLOSE)

    def send(self, dst_mode: int, size_src: int, tag: int, port: int = 0, f0: int = NONE, f1: int = NONE) -> None:
        self._emit(OP_SEND, dst_mode, size_src, tag, (port & 0xFFFF) | (f0 << 16) | (f1 << 24))

    def recv(self, handler: Optional[Target] = None) -> None:
        self._emit(OP_RECV, imm=("hi16", self._t(handler), 0))

    def recv_filter(self, rd: int, filter_index: int, asleep: Target, handler: Optional[Target] = None) -> None:
        """Fused ``while True: with ws.awake(): packet = socket.recv()`` loop head that also consumes the packets
        the program's own dispatch would ignore (see ``Simulator.packet_filter``)."""
        self._emit(OP_RECV_FILTER, rd, filter_index, imm=("hi16", self._t(handler), self._t(asleep)))

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
        raise TypeError(v)

    def assemble(self) -> Tuple[np.ndarray, np.ndarray]:
        """Returns (code uint64[n], entries (pc, prog) uint32[n_entries, 2])."""
        code = np.zeros(len(self._ins), dtype=np.uint64)
        for pc, (op, a, b, c, imm) in enumerate(self._ins):
            word = (op & 0xFF) | ((a & 0xFF) << 8) | ((b & 0xFF) << 16) | ((c & 0xFF) << 24) \
                | ((self._resolve(imm) & 0xFFFFFFFF) << 32)
            code[pc] = word
        entries = np.zeros((len(self.entries), 2), dtype=np.uint32)
        for i, (name, prog) in enumerate(self.entries):
            entries[i] = (self._labels[name], prog)
        return code, entries
