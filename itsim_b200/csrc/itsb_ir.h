/*
 * itsb_ir.h -- process-body instruction set interpreted by the engine (mirrored by itsim_b200/ir.py).
 *
 * ITsim process bodies are arbitrary Python functions that block deep inside plain calls; each
 * supported body is restated on the host as a small state machine over the primitives the reference's bodies use
 * (SURVEY.md section 8-a "Process-body IR"): advance, open/close socket, send/broadcast, recv, workstation
 * wake/sleep signal, spawn, throw, per-process remembered-packet list, variate draws.
 *
 * Instruction word (uint64): op[0:8] a[8:16] b[16:24] c[24:32] imm[32:64].
 * Blocking instructions carry the pc of their exception handler in imm[16:32] (0 = none: an Interrupt-class
 * exception ends the process silently, any other exception aborts the replica like it aborts Simulator.run).
 */
#ifndef ITSB_IR_H
#define ITSB_IR_H

enum {
    OP_NOP = 0,
    OP_EXIT = 1,        /* body returns                                                                        */
    OP_JMP = 2,         /* imm = target                                                                        */
    OP_LI = 3,          /* r[a] = imm                                                                          */
    OP_MOV = 4,         /* r[a] = src(b)                                                                       */
    OP_ADDI = 5,        /* r[a] = src(b) + (int32)imm                                                          */
    OP_JEQ = 6,         /* if src(a) == src(b) goto imm                                                        */
    OP_JNE = 7,
    OP_JEQI = 8,        /* if src(a) == imm[16:32] goto imm[0:16]                                              */
    OP_JNEI = 9,
    OP_DRAWI = 10,      /* r[a] = (int) next(dist imm)                                                         */
    OP_DRAW_ADDI = 11,  /* r[a] = int((double) src(b) + next(dist imm))       network_simple.py:258            */
    OP_ADVANCE = 12,    /* advance(next(dist imm[0:16]))                                                       */
    OP_ADVANCE_C = 13,  /* advance(consts[imm[0:16]])                                                          */
    OP_OPEN = 14,       /* open_socket(port imm[0:16]; 0 = next free unprivileged port)  overrides.py:337-369   */
    OP_CLOSE = 15,      /* leave the socket's with-block                                 overrides.py:366-369   */
    OP_SEND = 16,       /* a = dest mode, b = size src, c = tag, imm = port | f0src<<16 | f1src<<24             */
    OP_RECV = 17,       /* socket.recv()                                                 overrides.py:135-147   */
    OP_WAIT_AWAKE = 18, /* ws.wait_until_awake(); r[a] = wake epoch (a = 0xFF: discard)  network_simple.py:62-76 */
    OP_JASLEEP = 19,    /* if wake epoch != src(a) goto imm  (FellAsleep)                network_simple.py:75-76 */
    OP_NODE_SLEEP = 20, /* ws.sleep()                                                    network_simple.py:55-57 */
    OP_NODE_WAKE = 21,  /* ws.wake_up()                                                  network_simple.py:51-53 */
    OP_SPAWN = 22,      /* add(entry imm) on the same node, child regs = parent regs; r[a] = child handle        */
    OP_THROW = 23,      /* schedule(0, proc(src a).throw, exc imm)                       network_simple.py:37-38 */
    OP_LIST_CLEAR = 24, /* queries = []                                                  network_simple.py:180   */
    OP_LIST_PUSH = 25,  /* queries.append(packet)                                        network_simple.py:193   */
    OP_LIST_TAKE = 26,  /* first entry with f0 == src(a): load ENT_*, delete it; else goto imm   :203-211        */
    OP_STAT = 27,       /* user counter imm += 1                                                                 */
    OP_RECV_FILTER = 28,/* fused daemon loop head: "while True: with ws.awake(): packet = socket.recv(); if the
                           packet is not for this program: continue".  r[a] = wake epoch, b = filter index,
                           imm[0:16] = FellAsleep target, imm[16:32] = handler          network_simple.py:183-186 */
    OP_COUNT
};

/* operand sources */
enum {
    SRC_R0 = 0, SRC_R1 = 1, SRC_R2 = 2, SRC_R3 = 3,
    SRC_PKT_SRC_IP = 0x10, SRC_PKT_SRC_PORT = 0x11, SRC_PKT_SIZE = 0x12, SRC_PKT_TAG = 0x13,
    SRC_PKT_F0 = 0x14, SRC_PKT_F1 = 0x15,
    SRC_ENT_IP = 0x18, SRC_ENT_PORT = 0x19,
    SRC_NODE_ADDR = 0x20, SRC_NODE_INDEX = 0x21, SRC_NODE_HOST = 0x22, SRC_NODE_EPOCH = 0x23,
    SRC_NODE_BCAST = 0x24, SRC_SELF = 0x25, SRC_EXC = 0x26,
    SRC_ZERO = 0x30,
    SRC_NONE = 0xFF
};

/* OP_SEND destination modes */
enum {
    DST_BCAST = 0,      /* (broadcast address of the socket's network, port imm)     Socket.broadcast           */
    DST_PKT_SRC = 1,    /* packet.source of the packet just received                                            */
    DST_ENT = 2,        /* source remembered in the list entry just taken                                       */
    DST_SELF = 3,       /* (own default address, port imm)                                                      */
    DST_REGS = 4        /* (r2, r3 & 0xFFFF)                                                                    */
};

/* exception codes carried by THROW events; bit 7 = Interrupt class (greensim.Interrupt subclasses)             */
#define EXC_INTERRUPT_CLASS 0x80

#endif
