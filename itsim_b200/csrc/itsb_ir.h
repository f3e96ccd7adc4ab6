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
    /* ---- shipped-itsim primitives (itsim/machine, itsim/network/link.py) ---------------------------------------- */
    OP_BIND = 29,       /* Node.bind(port imm[0:16]; 0 = next ephemeral port 32768..60999)       node.py:168-200    */
    OP_SCLOSE = 30,     /* Socket.close(): free the port, turn the close signal on               socket.py:76-82    */
    OP_SENDL = 31,      /* Socket.send through routes and Link._transfer_packet; operands as OP_SEND
                                                                     socket.py:91-109; node.py:212-230; link.py:70-86 */
    OP_RECVT = 32,      /* Socket.recv(timeout = consts[imm[0:16]], 0xFFFF = None): greensim.select on the packet
                           and close signals with its helper processes                          socket.py:121-145   */
    OP_ADVANCE_X = 33,  /* advance(delay the thread was started with)                            thread.py:47        */
    OP_THREAD_EXIT = 34,/* exit_f bookkeeping; a = 1: then re-raise the pending exception        thread.py:49-50,62-70 */
    OP_CLONE = 35,      /* r[a] = thread.clone_in(consts[imm[16:32]], entry imm[0:16])           thread.py:35-39     */
    OP_FORK = 36,       /* r[a] = process.fork_exec_in(consts[imm[16:32]], entry imm[0:16])      process.py:87-96    */
    OP_JOIN = 37,       /* Thread.join(timeout consts[imm[0:16]] | None) on src(a)               thread.py:100-106   */
    OP_PWAIT = 38,      /* Process.wait(timeout) on src(a)                                       process.py:84-85    */
    OP_TKILL = 39,      /* Thread.kill() on src(a)                                               thread.py:87-98     */
    OP_PKILL = 40,      /* Process.kill() on src(a)                                              process.py:70-76    */
    OP_MARK = 41,       /* trace-only record of a model side effect (the tests' cemetary.append(name)): value imm, or
                           src(a) when a != 0xFF                                                                     */
    OP_SIG_WAIT = 42,   /* Signal.wait(timeout) on signal src(a)   (used by the select / balk helper programs)      */
    OP_SIG_ON = 43,     /* Signal.turn_on() on signal src(a)                                                         */
    OP_GSET = 44,       /* model variable g[imm & 7] = src(a)    (state shared between bodies, e.g. a dict of children) */
    OP_NOW_SUB = 45,    /* x[2:4] (f64) = consts[imm] - now      (advance(YEAR - now()))   then OP_ADVANCE_X2        */
    OP_ADVANCE_X2 = 46, /* advance(f64 in x[2:4])                                                                    */
    /* ---- small generic helpers the protocol bodies need (DHCP: dhcp/client.py, dhcp/server.py) ------------------- */
    OP_TLOAD = 47,      /* r[a] = table[imm + src(b)]        per-replica model table (dicts / sets of a daemon)       */
    OP_TSTORE = 48,     /* table[imm + src(b)] = src(a)                                                               */
    OP_ADD = 49,        /* r[a] = src(b) + src(c)                                                                     */
    OP_SUB = 50,        /* r[a] = src(b) - src(c)                                                                     */
    OP_JLTU = 51,       /* if src(a) < src(b) (unsigned) goto imm                                                     */
    OP_XNOW = 52,       /* x[0:2] = now()                    time_before_recv = now()            client.py:89         */
    OP_XSET = 53,       /* x[2:4] = consts[imm]              time_remaining = reservation_time   client.py:86         */
    OP_XELAPSE = 54,    /* x[2:4] -= now() - x[0:2]          time_remaining -= now() - before    client.py:93         */
    OP_JXLE0 = 55,      /* if x[2:4] <= 0.0 goto imm         while time_remaining > 0.0          client.py:87         */
    OP_RECVT_X = 56,    /* Socket.recv(timeout = x[2:4])                                         client.py:90         */
    OP_SETADDR = 57,    /* interface.address = src(a), re-rooted in the link's CIDR (first non-loopback interface)
                                                                                  interface.py:50-57; client.py:147 */
    OP_GETADDR = 58,    /* r[a] = interface.address                                              client.py:144        */
    OP_FORWARD = 59,    /* send the packet just received on, from (address of node imm[0:16], port imm[16:32]) to
                           (r2, r3): the one-way router of malware_simple.py:96-107                                  */
    OP_GETPORT = 60,    /* r[a] = socket.port                                                    socket.py:44-51      */
    OP_RECV_LOG = 61,   /* socket.recv() of a model that keeps connection records: OP_RECV followed by
                           correlate_inbound_packet                                              overrides.py:135-147   */
    OP_OPEN_AT = 62,    /* open_socket((address of another attachment, port)): imm = node | port << 16; the
                           process's current socket is unchanged                          malware_simple.py:98-99 */
    OP_SEND_LOG = 63,   /* OP_SEND followed by correlate_outbound_packet                         overrides.py:122-125   */
    OP_FORWARD_LOG = 64,/* OP_FORWARD followed by the same on the sending attachment's socket                           */
    OP_CLOSE_LOG = 65,  /* flush_logs, then OP_CLOSE                                              overrides.py:366-367   */
    OP_STAT_ADD = 66,   /* user counter imm += src(a)                                                              */
    OP_STAT_FIRST = 67, /* user counter imm = now in microseconds, the first time only (a time-to-first statistic) */
    OP_RECV_FILTER_LOG = 68, /* OP_RECV_FILTER with correlate_inbound_packet on every packet the loop head pops            */
    OP_COUNT
};

/* hidden programs appended by the assembler: greensim's own helper processes */
#define PROG_HIDDEN_BALK     0xF1   /* Queue.join timeout helper            */
#define PROG_HIDDEN_WAIT_ONE 0xF2   /* select() helper, one per signal      */

/* operand sources */
enum {
    SRC_R0 = 0, SRC_R1 = 1, SRC_R2 = 2, SRC_R3 = 3,
    SRC_PKT_SRC_IP = 0x10, SRC_PKT_SRC_PORT = 0x11, SRC_PKT_SIZE = 0x12, SRC_PKT_TAG = 0x13,
    SRC_PKT_F0 = 0x14, SRC_PKT_F1 = 0x15, SRC_PKT_DST_IP = 0x16,
    SRC_ENT_IP = 0x18, SRC_ENT_PORT = 0x19,
    SRC_NODE_ADDR = 0x20, SRC_NODE_INDEX = 0x21, SRC_NODE_HOST = 0x22, SRC_NODE_EPOCH = 0x23,
    SRC_NODE_BCAST = 0x24, SRC_SELF = 0x25, SRC_EXC = 0x26,
    SRC_SELF_THREAD = 0x27, SRC_SELF_TASK = 0x28, SRC_PID = 0x29,
    SRC_TAGS = 0x2A,    /* tags of the running process: what a process it creates inherits */
    SRC_ZERO = 0x30,
    SRC_G0 = 0x31,      /* model variables g0..g7 = 0x31..0x38 */
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

/* exception codes carried by THROW events; bit 7 = Interrupt class (greensim.Interrupt subclasses), bit 6 = hidden
 * (interrupts greensim sends to its own helper processes: not simulated events)                                 */
#define EXC_INTERRUPT_CLASS 0x80
#define EXC_HIDDEN          0x40
#define EXC_ITSIM_TIMEOUT   0x02   /* itsim.types.Timeout (itsim/types.py:220-224): a plain Exception             */
#define EXC_SOCKET_CLOSED   0x03   /* ValueError("Socket is closed")                                              */
#define EXC_THREAD_KILLED   0x81   /* ThreadKilled(Interrupt)            thread.py:10-11                          */
#define EXC_GS_TIMEOUT      0x82   /* greensim.Timeout(Interrupt), sent by the balk helper                        */
#define EXC_CLEANUP         0xC1   /* select()'s CleanUp                                                          */
#define EXC_CANCEL_BALK     0xC2   /* Queue.join's CancelBalk                                                     */

#endif
