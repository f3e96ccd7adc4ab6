/*
 * itsb_machine.cuh -- the shipped-`itsim` primitives (included by itsb_core.cuh; all out of line).
 *
 * What it restates (reference, /root/reference):
 *   generic Signal / Queue.join with timeout ("balk" helper process)   greensim, rows a3-a4
 *   Event (thread / process death)                                     itsim/simulator.py:49-77
 *   Link._transfer_packet: recipients, draw latency then bandwidth, lat + 8*size/bw      itsim/network/link.py:70-86
 *   Node._solve_transfer / _send_packet / _receive_packet                                itsim/machine/node.py:212-256
 *   Node.bind, ephemeral ports 32768..60999                                              itsim/machine/node.py:168-210
 *   Socket.send / recv(timeout) through greensim.select / close                          itsim/machine/socket.py:76-145
 *   Thread.run_in / clone_in / exit_f / kill / join                                      process_management/thread.py
 *   Process.exc_in / thread_complete / _die / kill / wait / fork_exec_in                 process_management/process.py
 *   Node.run_proc_in / next_proc_number                                                  itsim/machine/node.py:279-303
 *
 * greensim's helper processes (one per signal in select(), one per timed Queue.join) are real engine processes
 * running two tiny hidden programs the assembler appends (PROG_HIDDEN_*): their hops consume insertion-counter
 * values and queue positions exactly as in the reference, which is what makes same-timestamp ordering (the
 * integer-time known answers of tests/integ/test_thread_lifecycle.py) come out identical.  Hidden hops are not
 * counted as simulated events.
 */
#pragma once

namespace itsb {

ITSB_FN bool is_hidden_prog(u32 prog) { return prog >= 0xF0; }

/* ---- handles: index | generation << 16; a stale handle denotes a finished object --------------------------------- */
ITSB_FN u32 mk_handle(u32 idx, u32 gen) { return (idx & 0xFFFF) | (gen << 16); }
ITSB_FN u32 h_idx(u32 h) { return h & 0xFFFF; }
ITSB_FN u32 h_gen(u32 h) { return h >> 16; }

ITSB_FN bool proc_alive(Rep& R, u32 h) {
    const u32 p = h_idx(h);
    if (h == NIL32 || p >= R.hdr->L.max_proc) return false;
    const Pcb& P = R.pcb[p];
    return P.state != PS_FREE && P.gen == (h_gen(h) & GEN_MASK);
}

/* ---- signals ---------------------------------------------------------------------------------------------------------- */
ITSB_FN u32 sig_new(Rep& R) {
    Hdr* H = R.hdr;
    u32 top = H->sig_free_top;
    if (top == 0) { fail(H, ITSB_ERR_CAP_THREADS, 1); return NIL16; }
    top -= 1;
    H->sig_free_top = (u16)top;
    u32 s = R.sig_free[top];
    Sig& S = R.sig[s];
    S.head = NIL16; S.tail = NIL16; S.on = 0;
    return s;
}

ITSB_FN void sig_free(Rep& R, u32 s) {
    if (s == NIL16) return;
    Hdr* H = R.hdr;
    R.sig[s].gen = (u16)(R.sig[s].gen + 1);
    u32 top = H->sig_free_top;
    R.sig_free[top] = (u16)s;
    H->sig_free_top = (u16)(top + 1);
}

/* Signal.turn_on(): on, then resume every waiter in join order */
ITSB_FN void sig_on(Rep& R, u32 s) {
    if (s == NIL16) return;
    Sig& S = R.sig[s];
    S.on = 1;
    waiters_resume_all(R, S.head, S.tail);
}

/* Queue.join(timeout) half: enqueue, start the balk helper if the wait is timed, park.  (greensim; Appendix A) */
ITSB_FN void sig_join(Rep& R, u32 p, u32 s, bool timed, double timeout) {
    Hdr* H = R.hdr;
    Sig& S = R.sig[s];
    Pcb& P = R.pcb[p];
    waiters_append(R, S.head, S.tail, p, (u16)(WS_SIG | s));
    PcbX& X = R.pcbx[p];
    X.balker_h = NIL32;
    if (timed) {
        /* balker = add(balk, me): hidden process, r0 = my handle, its delay = the timeout */
        u32 regs[4] = {handle_of(R, p), 0, 0, 0};
        int b = proc_spawn(H, H->hidden_entry_balk, P.node, regs, 0.0);
        if (b >= 0) {
            const u64 tbits = dbits(timeout);
            R.pcbx[b].x[0] = (u32)tbits; R.pcbx[b].x[1] = (u32)(tbits >> 32);
            R.pcbx[b].thread = NIL16;
            X.balker_h = handle_of(R, (u32)b);
        }
    }
    P.state = PS_WAIT;
}

/* the `finally` of Queue.join: if the balk helper is still there, interrupt it (CancelBalk) */
ITSB_FN void cancel_balker(Rep& R, u32 p) {
    PcbX& X = R.pcbx[p];
    if (X.balker_h != NIL32 && proc_alive(R, X.balker_h))
        ev_push_inl(R, R.hdr->now, ev_pack(ITSB_EV_THROW, h_idx(X.balker_h), h_gen(X.balker_h), EXC_CANCEL_BALK));
    X.balker_h = NIL32;
}

/* Signal.wait(timeout) as a re-executed instruction: returns true if the process parked.  P.phase = 1 marks "I was
 * parked in this wait", so the resume runs join()'s finally before looking at the signal again. */
ITSB_FN bool sig_wait(Rep& R, u32 p, u32 s, u32 timeout_const) {
    Pcb& P = R.pcb[p];
    if (P.phase == 1) { cancel_balker(R, p); P.phase = 0; }
    if (s == NIL16 || R.sig[s].on) return false;
    sig_join(R, p, s, timeout_const != 0xFFFF, timeout_const != 0xFFFF ? R.m.consts[timeout_const] : 0.0);
    P.phase = 1;
    return true;
}

/* ---- threads and processes (itsim.machine.process_management) ------------------------------------------------------------- */
ITSB_FN bool thr_live(Rep& R, u32 h) {
    const u32 t = h_idx(h);
    return h != NIL32 && t < R.hdr->L.max_thr && R.thr[t].gen == h_gen(h) && R.thr[t].live;
}
ITSB_FN bool task_live(Rep& R, u32 h) {
    const u32 t = h_idx(h);
    return h != NIL32 && t < R.hdr->L.max_task && R.task[t].gen == h_gen(h) && R.task[t].live;
}

/* Process(n, node, parent) */
ITSB_FN int task_new(Rep& R, u32 node, u32 parent_h) {
    Hdr* H = R.hdr;
    u32 top = H->task_free_top;
    if (top == 0) { fail(H, ITSB_ERR_CAP_THREADS, 2); return -1; }
    top -= 1;
    H->task_free_top = (u16)top;
    u32 t = R.task_free[top];
    Task& T = R.task[t];
    NodeDyn& N = R.node[node];
    T.pid = (u16)N.awake;            /* machine nodes reuse this word as _process_counter (node.py:291-293) */
    N.awake += 1;
    T.node = (u16)node; T.parent = parent_h; T.nthreads = 0; T.thread_counter = 0;
    T.thr_head = NIL16; T.thr_tail = NIL16; T.live = 1;
    T.dead_sig = (u16)sig_new(R);
    return (int)t;
}

/* Thread(sim, process, n) + bookkeeping of Process.exc_in (process.py:45-51) */
ITSB_FN int thr_new(Rep& R, u32 task) {
    Hdr* H = R.hdr;
    u32 top = H->thr_free_top;
    if (top == 0) { fail(H, ITSB_ERR_CAP_THREADS, 3); return -1; }
    top -= 1;
    H->thr_free_top = (u16)top;
    u32 t = R.thr_free[top];
    Thr& T = R.thr[t];
    Task& K = R.task[task];
    T.task = (u16)task; T.ncomp = 0; T.comp_head = NIL16; T.comp_tail = NIL16; T.live = 1;
    T.n = K.thread_counter; K.thread_counter += 1;
    T.dead_sig = (u16)sig_new(R);
    T.next = NIL16;
    if (K.thr_tail == NIL16) K.thr_head = (u16)t; else R.thr[K.thr_tail].next = (u16)t;
    K.thr_tail = (u16)t;
    K.nthreads += 1;
    return (int)t;
}

/* Thread.run_in(time, f): sim.add(wrap_computation, sim_comp, time) -- the computation starts now, advances `time`,
 * then runs the body (thread.py:41-57) */
ITSB_FN int thr_run_in(Rep& R, u32 thr, u32 entry, u32 delay_const, const u32* regs, u32 inherit_sock = NIL16) {
    Hdr* H = R.hdr;
    Thr& T = R.thr[thr];
    const u32 node = R.task[T.task].node;
    int p = proc_spawn(H, entry, node, regs, 0.0);
    if (p < 0) return -1;
    PcbX& X = R.pcbx[p];
    const u64 dbit = dbits(R.m.consts[delay_const]);
    X.x[0] = (u32)dbit; X.x[1] = (u32)(dbit >> 32); X.x[2] = 0; X.x[3] = 0;
    X.thread = (u16)thr; X.tnext = NIL16; X.balker_h = NIL32; X.sel_common = NIL16; X.sel_h1 = NIL32; X.sel_h2 = NIL32;
    if (T.comp_tail == NIL16) T.comp_head = (u16)p; else R.pcbx[T.comp_tail].tnext = (u16)p;
    T.comp_tail = (u16)p;
    T.ncomp += 1;
    R.pcb[p].sock = (u16)inherit_sock;      /* a body started with the caller's socket as argument (node.py:343-346) */
    return p;
}

/* Node.run_proc_in (node.py:279-283): new Process (pid from the node's counter), first Thread, first computation */
ITSB_COLD u32 m_run_proc_in(Hdr* H, u32 node, u32 parent_h, u32 entry, u32 delay_const, const u32* regs) {
    Rep R; rebind(R, H);
    int task = task_new(R, node, parent_h);
    if (task < 0) return NIL32;
    int thr = thr_new(R, (u32)task);
    if (thr < 0) return NIL32;
    thr_run_in(R, (u32)thr, entry, delay_const, regs);
    return mk_handle((u32)task, R.task[task].gen);
}

/* Process._die (process.py:78-82): fire the death event; parent / node bookkeeping has no simulated effect */
ITSB_FN void task_die(Rep& R, u32 t) {
    Hdr* H = R.hdr;
    Task& T = R.task[t];
    sig_on(R, T.dead_sig);
    sig_free(R, T.dead_sig);
    T.live = 0;
    T.gen = (u16)(T.gen + 1);
    u32 top = H->task_free_top;
    R.task_free[top] = (u16)t;
    H->task_free_top = (u16)(top + 1);
}

/* Thread.exit_f (thread.py:62-70): drop the computation; last one out -> process.thread_complete, then the thread's
 * own death event */
ITSB_COLD void m_thread_exit(Hdr* H, u32 p) {
    Rep R; rebind(R, H);
    PcbX& X = R.pcbx[p];
    const u32 thr = X.thread;
    if (thr == NIL16) return;
    Thr& T = R.thr[thr];
    /* _computations.remove(sim_comp) */
    u16 prev = NIL16, c = T.comp_head;
    while (c != NIL16 && c != p) { prev = c; c = R.pcbx[c].tnext; }
    if (c == p) {
        if (prev == NIL16) T.comp_head = X.tnext; else R.pcbx[prev].tnext = X.tnext;
        if (T.comp_tail == p) T.comp_tail = prev;
        T.ncomp -= 1;
    }
    X.thread = NIL16;
    if (T.ncomp != 0) return;
    /* process.thread_complete(self) (process.py:56-59) */
    Task& K = R.task[T.task];
    u16 tp = NIL16, tc = K.thr_head;
    while (tc != NIL16 && tc != thr) { tp = tc; tc = R.thr[tc].next; }
    if (tc == thr) {
        if (tp == NIL16) K.thr_head = T.next; else R.thr[tp].next = T.next;
        if (K.thr_tail == thr) K.thr_tail = tp;
        K.nthreads -= 1;
    }
    if (K.nthreads == 0) task_die(R, T.task);
    /* self._event_dead.fire() */
    sig_on(R, T.dead_sig);
    sig_free(R, T.dead_sig);
    T.live = 0;
    T.gen = (u16)(T.gen + 1);
    u32 top = H->thr_free_top;
    R.thr_free[top] = (u16)thr;
    H->thr_free_top = (u16)(top + 1);
}

/* Thread.kill (thread.py:87-98): ThreadKilled into every computation, in creation order */
ITSB_FN void thr_kill(Rep& R, u32 thr) {
    for (u16 c = R.thr[thr].comp_head; c != NIL16; c = R.pcbx[c].tnext)
        ev_push_inl(R, R.hdr->now, ev_pack(ITSB_EV_THROW, c, R.pcb[c].gen, EXC_THREAD_KILLED));
}

/* ---- network: routes, link transfer, delivery ------------------------------------------------------------------------------ */

/* Node._solve_transfer (node.py:212-225): longest prefix over interfaces (loopback first) x routes, first seen wins
 * ties.  Returns the interface index or -1 (NoRouteToHost); *hop = route.get_hop(dest). */
ITSB_FN int solve_transfer(Rep& R, u32 node, u32 dest, u32* hop) {
    const u32 fl = R.m.nodes[node].flags;
    const u32 first = fl & 0xFFFF, n = fl >> 16;
    int best_if = -1;
    u32 best_len = 0, best_hop = 0;
    for (u32 i = first; i < first + n; ++i) {
        const itsb_iface I = R.m.ifaces[i];
        for (u32 r = I.first_route; r < I.first_route + I.n_routes; ++r) {
            const itsb_route Q = R.m.routes[r];
            if ((dest & Q.mask) == Q.net && (best_if < 0 || Q.prefixlen > best_len)) {
                best_if = (int)i; best_len = Q.prefixlen;
                best_hop = (Q.kind == ITSB_ROUTE_RELAY) ? Q.gateway : dest;
            }
        }
    }
    *hop = best_hop;
    return best_if;
}

/* Socket.send -> Node._send_packet -> Link._transfer_packet (socket.py:91-109; node.py:227-230; link.py:70-86) */
ITSB_COLD void m_send(Hdr* H, u32 p, u32 dst_ip, u32 dst_port, u32 size, u32 tag, u32 f0, u32 f1) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    if (P.sock == NIL16 || R.sig[R.sock[P.sock].wait_tail].on) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return; }
    u32 hop;
    const int ifx = solve_transfer(R, P.node, dst_ip, &hop);
    if (ifx < 0) { fail(H, ITSB_EXC_NO_ROUTE, dst_ip); return; }
    const itsb_iface I = R.m.ifaces[ifx];
    const itsb_net NET = R.m.nets[I.net];
    /* recipients: every node of the link for the broadcast address, else the first node (connection order) one of
       whose interfaces carries the hop address */
    u32 rcpt = 0;   /* 0 = broadcast, else node + 1 */
    if (hop != NET.bcast) {
        bool found = false;
        for (u32 m = 0; m < NET.n_nodes && !found; ++m) {
            const u32 node = R.m.members[NET.first_node + m];
            const u32 fl = R.m.nodes[node].flags;
            for (u32 i = fl & 0xFFFF; i < (fl & 0xFFFF) + (fl >> 16); ++i)
                if (R.iface_addr[i] == hop) { rcpt = node + 1; found = true; break; }
        }
        if (!found) { fail(H, ITSB_EXC_NO_SUCH_ADDRESS, hop); return; }
    }
    const double lat = draw_dist(H, &R.m.dists[NET.lat_dist]);
    const double bw = draw_dist(H, &R.m.dists[NET.bw_dist]);
    const double duration = lat + (double)(8ull * (u64)size) / bw;
    if (duration < 0.0) { fail(H, ITSB_EXC_NEGATIVE_DELAY, P.pc); return; }
    u32 k = arena_alloc(R);
    if (k == NIL32) return;
    Pkt K;
    K.src_ip = R.iface_addr[ifx]; K.dst_ip = dst_ip;
    K.ports = (u32)R.sock[P.sock].port | (dst_port << 16);
    K.size = size; K.tag = (tag & 0xFFFF) | (rcpt << 16); K.f0 = f0; K.f1 = f1;
    K.meta = (u32)P.node | (I.net << 16);
    *pkt_at(R, k) = K;
    stat_add(R, ITSB_TM_PACKETS_SENT, 1);
    stat_add(R, ITSB_TM_BYTES_SENT, size);
    stat_add(R, ITSB_TM_LAT0 + latency_bin(duration), 1);
    const u32 count = rcpt ? 1u : NET.n_nodes;
    ev_push_inl(R, H->now + duration, ev_pack_pkt(ITSB_EV_DELIVER, k));   /* for node in recipients: add_in(...) */
    H->seq += count - 1;
}

/* Router forwarding: the packet (arena node k, left untouched) goes out again through the route to its destination;
 * errors of the lookup are the forwarding node's (NoRouteToHost / NoSuchAddress abort the run like any exception
 * raised inside _receive_packet). */
ITSB_COLD void m_forward(Hdr* H, u32 node, u32 k) {
    Rep R; rebind(R, H);
    const Pkt K = *pkt_at(R, k);
    u32 hop;
    const int ifx = solve_transfer(R, node, K.dst_ip, &hop);
    if (ifx < 0) { fail(H, ITSB_EXC_NO_ROUTE, K.dst_ip); return; }
    const itsb_iface I = R.m.ifaces[ifx];
    const itsb_net NET = R.m.nets[I.net];
    u32 rcpt = 0;
    if (hop != NET.bcast) {
        bool found = false;
        for (u32 m = 0; m < NET.n_nodes && !found; ++m) {
            const u32 cand = R.m.members[NET.first_node + m];
            const u32 fl = R.m.nodes[cand].flags;
            for (u32 i = fl & 0xFFFF; i < (fl & 0xFFFF) + (fl >> 16); ++i)
                if (R.iface_addr[i] == hop) { rcpt = cand + 1; found = true; break; }
        }
        if (!found) { fail(H, ITSB_EXC_NO_SUCH_ADDRESS, hop); return; }
    }
    const double lat = draw_dist(H, &R.m.dists[NET.lat_dist]);
    const double bw = draw_dist(H, &R.m.dists[NET.bw_dist]);
    const double duration = lat + (double)(8ull * (u64)K.size) / bw;
    u32 k2 = arena_alloc(R);
    if (k2 == NIL32) return;
    Pkt F = K;
    F.tag = (K.tag & 0xFFFF) | (rcpt << 16);
    F.meta = node | (I.net << 16);
    *pkt_at(R, k2) = F;
    stat_add(R, ITSB_TM_LAT0 + latency_bin(duration), 1);
    const u32 count = rcpt ? 1u : NET.n_nodes;
    ev_push_inl(R, H->now + duration, ev_pack_pkt(ITSB_EV_DELIVER, k2));
    H->seq += count - 1;
}

/* Node._receive_packet for every recipient (node.py:232-243), in connection order */
ITSB_COLD void m_deliver(Hdr* H, u32 k) {
    Rep R; rebind(R, H);
    const Pkt K = *pkt_at(R, k);
    const itsb_net NET = R.m.nets[K.meta >> 16];
    const u32 rcpt = K.tag >> 16;
    const u32 count = rcpt ? 1u : NET.n_nodes;
    const u32 dst_port = K.ports >> 16;
    u32 forward_from = NIL32;
    for (u32 m = 0; m < count; ++m) {
        const u32 node = rcpt ? rcpt - 1 : R.m.members[NET.first_node + m];
        const u32 fl = R.m.nodes[node].flags;
        bool mine = false;
        for (u32 i = fl & 0xFFFF; i < (fl & 0xFFFF) + (fl >> 16); ++i) {
            const u32 a = R.iface_addr[i];
            if (a == K.dst_ip || R.m.nets[R.m.ifaces[i].net].bcast == K.dst_ip) mine = true;
        }
        int s = -1;
        if (mine)
            for (u16 q = R.node[node].sock_head; q != NIL16; q = R.sock[q].next)
                if (R.sock[q].port == dst_port) { s = (int)q; break; }
        count_event(R, ITSB_EV_DELIVER, 10, node, (K.size & 0x7FFFFFFFu) | (s >= 0 ? 0x80000000u : 0u));
        if (!mine && (R.m.nodes[node].host & ITSB_NODE_FORWARDS)) {
            /* Node.handle_packet_transit overridden with the forwarding rule the reference documents as intent
               (sphinx/source/networks.rst:275-280):
               interface, hop = self._solve_transfer(dest); interface.link._transfer_packet(packet, hop).
               Only a unicast hop can end at a node the packet is not for, so this is the pass's only recipient
               and the forwarding itself runs after the loop. */
            forward_from = node;
            continue;
        }
        if (s >= 0) {
            /* Socket._enqueue: queue.put(packet); packet_signal.turn_on() */
            u32 an = arena_alloc(R);
            if (an == NIL32) return;
            ArenaNode A;
            A.next = NIL32; A.src_ip = K.src_ip; A.ports = K.ports; A.size = K.size;
            A.tag = K.tag & 0xFFFF; A.f0 = K.f0; A.f1 = K.f1; A.pad = K.dst_ip;
            R.arena[an] = A;
            Sock& S = R.sock[s];
            if (S.q_tail == NIL32) S.q_head = an; else R.arena[S.q_tail].next = an;
            S.q_tail = an;
            S.q_len += 1;
            sig_on(R, S.wait_head);
            stat_add(R, ITSB_TM_DELIVERED, 1);
        } else {
            stat_add(R, ITSB_TM_DROPPED, 1);
        }
    }
    if (forward_from != NIL32) m_forward(H, forward_from, k);
    arena_release(R, k);
}

/* Node.bind (node.py:178-200): explicit port or the next free port of the persistent 32768..60999 cycle; ports 0 and
 * 65535 are never free.  The socket's two signals start off (socket.py:35-36). */
/* releases the socket slot of a finishing process (queued packets, both signals) */
ITSB_FN void m_sock_release(Rep& R, u32 p);

ITSB_COLD int m_bind(Hdr* H, u32 p, u32 port) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    NodeDyn& N = R.node[P.node];
    if (P.sock != NIL16) m_sock_release(R, p);   /* the name is rebound: the old Socket is finalised (socket.py:59-68) */
    if (port == 0) {
        u32 c = N.port_cursor, visited = 0;
        for (;;) {
            u32 cand = c;
            c = (c + 1 >= 61000u) ? 32768u : c + 1u;
            if (!port_in_use(R, P.node, cand)) { port = cand; break; }
            visited += 1;
            if (visited >= 61000u - 32768u) { fail(H, ITSB_EXC_PORTS_EXHAUSTED, 0); return -1; }
        }
        N.port_cursor = (u16)c;
    }
    if (port == 0 || port == 65535u || port_in_use(R, P.node, port)) { fail(H, ITSB_EXC_PORT_IN_USE, port); return -1; }
    u32 top = H->sock_free_top;
    if (top == 0) { fail(H, ITSB_ERR_CAP_SOCKETS, port); return -1; }
    top -= 1;
    H->sock_free_top = (u16)top;
    u32 s = R.sock_free[top];
    Sock& S = R.sock[s];
    S.port = (u16)port; S.node = P.node; S.owner = (u16)p;
    S.wait_head = (u16)sig_new(R);   /* _packet_signal */
    S.wait_tail = (u16)sig_new(R);   /* _close_signal  */
    S.q_head = NIL32; S.q_tail = NIL32; S.q_len = 0;
    S.next = N.sock_head;
    N.sock_head = (u16)s;
    P.sock = (u16)s;
    {
        const u32 fl = R.m.nodes[P.node].flags;
        const u32 th = R.pcbx[p].thread;
        net_event_put(H, 1, P.node, port, R.iface_addr[(fl & 0xFFFF) + ((fl >> 16) > 1 ? 1 : 0)], P.prog,
                      th == NIL16 ? -1 : (int)R.task[R.thr[th].task].pid, 1);
    }
    return (int)s;
}

/* Socket.close (socket.py:76-82): the port is free again at once; whoever waits on the close signal wakes.  The slot
 * itself stays until its owner ends, so a receiver blocked on it still finds it (closed). */
ITSB_COLD void m_close(Hdr* H, u32 p) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    const u32 s = P.sock;
    if (s == NIL16) return;
    Sock& S = R.sock[s];
    if (R.sig[S.wait_tail].on) return;
    NodeDyn& N = R.node[S.node];
    {
        const u32 fl = R.m.nodes[S.node].flags;
        const u32 th = R.pcbx[p].thread;
        net_event_put(H, 2, S.node, S.port, R.iface_addr[(fl & 0xFFFF) + ((fl >> 16) > 1 ? 1 : 0)], P.prog,
                      th == NIL16 ? -1 : (int)R.task[R.thr[th].task].pid, 1);
    }
    if (N.sock_head == s) N.sock_head = S.next;
    else {
        u16 q = N.sock_head;
        while (q != NIL16 && R.sock[q].next != s) q = R.sock[q].next;
        if (q != NIL16) R.sock[q].next = S.next;
    }
    sig_on(R, S.wait_tail);
}

/* releases the socket slot of a finishing process (queued packets, both signals) */
ITSB_FN void m_sock_release(Rep& R, u32 p) {
    Hdr* H = R.hdr;
    Pcb& P = R.pcb[p];
    const u32 s = P.sock;
    if (s == NIL16 || R.sock[s].owner != p) { P.sock = NIL16; return; }
    Sock& S = R.sock[s];
    if (!R.sig[S.wait_tail].on) {   /* never closed: the reference's weak reference frees the port on collection */
        NodeDyn& N = R.node[S.node];
        if (N.sock_head == s) N.sock_head = S.next;
        else {
            u16 q = N.sock_head;
            while (q != NIL16 && R.sock[q].next != s) q = R.sock[q].next;
            if (q != NIL16) R.sock[q].next = S.next;
        }
    }
    u32 n = S.q_head;
    while (n != NIL32) { u32 nx = R.arena[n].next; arena_release(R, n); n = nx; }
    sig_free(R, S.wait_head);
    sig_free(R, S.wait_tail);
    S.q_head = NIL32; S.q_tail = NIL32; S.q_len = 0; S.port = 0;
    u32 top = H->sock_free_top;
    R.sock_free[top] = (u16)s;
    H->sock_free_top = (u16)(top + 1);
    P.sock = NIL16;
}

/* select()'s `finally`: interrupt both helpers (the interrupt is scheduled whether or not the helper is still there:
 * a finished one makes a dead hop that still takes a counter value), drop the private signal */
ITSB_FN void select_cleanup(Rep& R, u32 p) {
    Hdr* H = R.hdr;
    PcbX& X = R.pcbx[p];
    const u32 hs[2] = {X.sel_h1, X.sel_h2};
    for (int i = 0; i < 2; ++i) {
        if (hs[i] == NIL32) continue;
        if (proc_alive(R, hs[i])) ev_push_inl(R, H->now, ev_pack(ITSB_EV_THROW, h_idx(hs[i]), h_gen(hs[i]), EXC_CLEANUP));
        else H->seq += 1;
    }
    sig_free(R, X.sel_common);
    X.sel_common = NIL16; X.sel_h1 = NIL32; X.sel_h2 = NIL32;
}

/* The exception (if any) an instruction raises synchronously goes to its own handler like one thrown from outside. */
ITSB_FN int raise_here(Rep& R, u32 p, u32 exc) {
    Hdr* H = R.hdr;
    Pcb& P = R.pcb[p];
    const u32 handler = (u32)(R.m.code[P.pc] >> 48);
    P.exc = (u16)exc;
    H->vol[SRC_EXC - 0x10] = exc;
    P.phase = 0;
    if (handler == 0) {
        if (!(exc & EXC_INTERRUPT_CLASS)) { fail(H, ITSB_EXC_UNCAUGHT, ((u64)exc << 32) | P.pc); return XR_STOP; }
        P.pc = 0;
        return XR_CONTINUE;
    }
    P.pc = (u16)handler;
    return XR_CONTINUE;
}

/* Instructions 29.. (shipped-itsim primitives). */
ITSB_COLD int exec_machine(Hdr* H, u32 p, u64 ins) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    PcbX& X = R.pcbx[p];
    const u32 op = (u32)(ins & 0xFF), a = (u32)((ins >> 8) & 0xFF), b = (u32)((ins >> 16) & 0xFF),
              c = (u32)((ins >> 24) & 0xFF), imm = (u32)(ins >> 32);
    /* the thread / signal / link instructions need their tables; the small generic helpers do not */
    const bool generic = op == OP_MARK || op == OP_GETPORT || op == OP_GSET || op == OP_TLOAD || op == OP_TSTORE || op == OP_ADD ||
                         op == OP_SUB || op == OP_JLTU;
    if (H->L.max_sig == 0 && !generic) { fail(H, ITSB_ERR_BAD_OPCODE, P.pc); return XR_STOP; }
    switch (op) {
    case OP_BIND:
        if (m_bind(H, p, imm & 0xFFFF) < 0) return XR_STOP;
        if (a == 1) R.sock[P.sock].owner = NIL16;     /* detached: outlives the binder (networking daemon) */
        P.pc += 1;
        return XR_CONTINUE;
    case OP_SCLOSE:
        m_close(H, p);
        P.pc += 1;
        return XR_CONTINUE;
    case OP_SENDL: {
        const u32* v = H->vol;
        u32 dst_ip, dst_port = imm & 0xFFFF;
        switch (a) {
        case DST_BCAST: {   /* broadcast address of the node's first non-loopback interface */
            const u32 fl = R.m.nodes[P.node].flags;
            dst_ip = R.m.nets[R.m.ifaces[(fl & 0xFFFF) + ((fl >> 16) > 1 ? 1 : 0)].net].bcast;
            break;
        }
        case DST_PKT_SRC: dst_ip = v[SRC_PKT_SRC_IP - 0x10]; dst_port = v[SRC_PKT_SRC_PORT - 0x10]; break;
        default: dst_ip = P.r[2]; dst_port = P.r[3] & 0xFFFF; break;
        }
        if (P.sock != NIL16 && R.sig[R.sock[P.sock].wait_tail].on) return raise_here(R, p, EXC_SOCKET_CLOSED);
        u32 f0s = (imm >> 16) & 0xFF, f1s = (imm >> 24) & 0xFF;
        m_send(H, p, dst_ip, dst_port, src_val(R, P, b), c,
               f0s == SRC_NONE ? 0 : src_val(R, P, f0s), f1s == SRC_NONE ? 0 : src_val(R, P, f1s));
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_RECVT:
    case OP_RECVT_X: {
        if (P.sock == NIL16) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return XR_STOP; }
        Sock& S = R.sock[P.sock];
        bool timed;
        double timeout;
        if (op == OP_RECVT) { timed = (imm & 0xFFFF) != 0xFFFF; timeout = timed ? R.m.consts[imm & 0xFFFF] : 0.0; }
        else { timed = true; timeout = bitsd(((u64)X.x[3] << 32) | X.x[2]); }
        if (P.phase == 0) {
            if (R.sig[S.wait_tail].on) return raise_here(R, p, EXC_SOCKET_CLOSED);
            /* greensim.select(packet_signal, close_signal, timeout=...) */
            const u32 common = sig_new(R);
            if (common == NIL16) return XR_STOP;
            X.sel_common = (u16)common;
            u32 regs[4] = {S.wait_head, common, 0, 0};
            int h1 = proc_spawn(H, H->hidden_entry_wait_one, P.node, regs, 0.0);
            regs[0] = S.wait_tail;
            int h2 = proc_spawn(H, H->hidden_entry_wait_one, P.node, regs, 0.0);
            if (h1 < 0 || h2 < 0) return XR_STOP;
            R.pcbx[h1].thread = NIL16; R.pcbx[h2].thread = NIL16;
            X.sel_h1 = handle_of(R, (u32)h1); X.sel_h2 = handle_of(R, (u32)h2);
            sig_join(R, p, common, timed, timeout);   /* common.wait(timeout): just created, so it is off */
            P.phase = 1;
            return XR_STOP;
        }
        /* resumed: join()'s finally, then `while not common.is_on` */
        cancel_balker(R, p);
        if (!R.sig[X.sel_common].on) {
            sig_join(R, p, X.sel_common, timed, timeout);
            return XR_STOP;
        }
        select_cleanup(R, p);
        P.phase = 0;
        if (R.sig[S.wait_tail].on) return raise_here(R, p, EXC_SOCKET_CLOSED);
        if (S.q_len == 0) { fail(H, ITSB_ERR_STATE, P.pc); return XR_STOP; }   /* Queue.get() would block forever */
        const ArenaNode A = sock_pop(R, S);
        set_pkt_regs(R, A);
        if (S.q_len == 0) R.sig[S.wait_head].on = 0;      /* packet_signal.turn_off() */
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_ADVANCE_X:
    case OP_ADVANCE_X2: {
        const u32 o = (op == OP_ADVANCE_X) ? 0 : 2;
        const double d = bitsd(((u64)X.x[o + 1] << 32) | X.x[o]);
        if (d < 0.0) { fail(H, ITSB_EXC_NEGATIVE_DELAY, P.pc); return XR_STOP; }
        P.timer_tok = (u16)(P.timer_tok + 1);
        int slot = ev_push_inl(R, H->now + d, ev_pack(ITSB_EV_TIMER, p, P.gen, P.timer_tok));
        P.timer_ev = slot < 0 ? NIL16 : (u16)slot;
        P.state = PS_TIMER;
        return XR_STOP;
    }
    case OP_NOW_SUB: {
        const double d = R.m.consts[imm & 0xFFFF] - H->now;
        const u64 bits = dbits(d);
        X.x[2] = (u32)bits; X.x[3] = (u32)(bits >> 32);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_THREAD_EXIT: {
        m_thread_exit(H, p);
        m_sock_release(R, p);
        if (a == 1 && !(P.exc & EXC_INTERRUPT_CLASS)) { fail(H, ITSB_EXC_UNCAUGHT, ((u64)P.exc << 32) | P.pc); return XR_STOP; }
        proc_exit(H, p);
        return XR_STOP;
    }
    case OP_CLONE: {
        if (X.thread == NIL16) { fail(H, ITSB_ERR_STATE, P.pc); return XR_STOP; }
        int t = thr_new(R, R.thr[X.thread].task);
        if (t < 0) return XR_STOP;
        const u32 gen = R.thr[t].gen;
        if (thr_run_in(R, (u32)t, imm & 0xFFFF, imm >> 16, P.r, P.sock) < 0) return XR_STOP;
        if (a != SRC_NONE) P.r[a & 3] = mk_handle((u32)t, gen);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_FORK: {
        if (X.thread == NIL16) { fail(H, ITSB_ERR_STATE, P.pc); return XR_STOP; }
        const u32 me = R.thr[X.thread].task;
        u32 h = m_run_proc_in(H, P.node, mk_handle(me, R.task[me].gen), imm & 0xFFFF, imm >> 16, P.r);
        if (h == NIL32) return XR_STOP;
        if (a != SRC_NONE) P.r[a & 3] = h;
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_JOIN:
    case OP_PWAIT: {
        /* Event.wait(timeout): signal.wait(timeout), greensim.Timeout -> itsim Timeout (simulator.py:70-77) */
        const u32 h = src_val(R, P, a);
        u32 sig = NIL16;
        if (op == OP_JOIN) { if (thr_live(R, h)) sig = R.thr[h_idx(h)].dead_sig; }
        else { if (task_live(R, h)) sig = R.task[h_idx(h)].dead_sig; }
        if (sig_wait(R, p, sig, imm & 0xFFFF)) return XR_STOP;
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_TKILL: {
        const u32 h = src_val(R, P, a);
        if (thr_live(R, h)) thr_kill(R, h_idx(h));
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_PKILL: {
        /* Process.kill (process.py:70-76): every thread, in creation order (a process without threads is already
           dead here) */
        const u32 h = src_val(R, P, a);
        if (task_live(R, h))
            for (u16 t = R.task[h_idx(h)].thr_head; t != NIL16; t = R.thr[t].next) thr_kill(R, t);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_MARK:
        if (R.trace) trace_put(H, ITSB_TRACE_MARK, 0, 0, a == SRC_NONE ? imm : src_val(R, P, a));
        P.pc += 1;
        return XR_CONTINUE;
    case OP_GETPORT:
        P.r[a & 3] = P.sock == NIL16 ? 0 : R.sock[P.sock].port;
        P.pc += 1;
        return XR_CONTINUE;
    case OP_SIG_WAIT:
        if (sig_wait(R, p, src_val(R, P, a) & 0xFFFF, imm & 0xFFFF)) return XR_STOP;
        P.pc += 1;
        return XR_CONTINUE;
    case OP_SIG_ON:
        sig_on(R, src_val(R, P, a) & 0xFFFF);
        P.pc += 1;
        return XR_CONTINUE;
    case OP_GSET:
        H->vol[SRC_G0 - 0x10 + (imm & 7)] = src_val(R, P, a);
        P.pc += 1;
        return XR_CONTINUE;
    case OP_TLOAD:
    case OP_TSTORE: {
        const u32 idx = imm + src_val(R, P, b);
        if (idx >= H->L.table_words) { fail(H, ITSB_ERR_BAD_OPCODE, P.pc); return XR_STOP; }
        if (op == OP_TLOAD) P.r[a & 3] = R.tab[idx]; else R.tab[idx] = src_val(R, P, a);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_ADD:
        P.r[a & 3] = src_val(R, P, b) + src_val(R, P, c); P.pc += 1;
        return XR_CONTINUE;
    case OP_SUB:
        P.r[a & 3] = src_val(R, P, b) - src_val(R, P, c); P.pc += 1;
        return XR_CONTINUE;
    case OP_JLTU:
        P.pc = (src_val(R, P, a) < src_val(R, P, b)) ? (u16)imm : (u16)(P.pc + 1);
        return XR_CONTINUE;
    case OP_XNOW: {
        const u64 bits = dbits(H->now);
        X.x[0] = (u32)bits; X.x[1] = (u32)(bits >> 32);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_XSET: {
        const u64 bits = dbits(R.m.consts[imm & 0xFFFF]);
        X.x[2] = (u32)bits; X.x[3] = (u32)(bits >> 32);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_XELAPSE: {
        const double before = bitsd(((u64)X.x[1] << 32) | X.x[0]);
        const double remaining = bitsd(((u64)X.x[3] << 32) | X.x[2]) - (H->now - before);
        const u64 bits = dbits(remaining);
        X.x[2] = (u32)bits; X.x[3] = (u32)(bits >> 32);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_JXLE0: {
        const double remaining = bitsd(((u64)X.x[3] << 32) | X.x[2]);
        P.pc = !(remaining > 0.0) ? (u16)imm : (u16)(P.pc + 1);
        return XR_CONTINUE;
    }
    case OP_SETADDR:
    case OP_GETADDR: {
        const u32 fl = R.m.nodes[P.node].flags;
        const u32 ifx = (fl & 0xFFFF) + ((fl >> 16) > 1 ? 1 : 0);
        if (op == OP_GETADDR) P.r[a & 3] = R.iface_addr[ifx];
        else {
            const itsb_net NET = R.m.nets[R.m.ifaces[ifx].net];
            R.iface_addr[ifx] = NET.cidr_net + (src_val(R, P, a) & ~NET.cidr_mask);   /* as_address(ar, cidr) */
        }
        P.pc += 1;
        return XR_CONTINUE;
    }
    default:
        fail(H, ITSB_ERR_BAD_OPCODE, P.pc);
        return XR_STOP;
    }
}

/* The part of greenlet.throw that depends on what the target was parked in, for shipped-itsim waits: leave the
 * signal's queue, run join()'s and select()'s `finally`, translate greensim.Timeout (socket.py:135-136;
 * simulator.py:76-77). */
ITSB_FN u32 machine_unwind(Rep& R, u32 p, u32 exc) {
    Pcb& P = R.pcb[p];
    if (R.hdr->L.max_sig == 0) return exc;
    if ((P.wsig & 0xC000) == WS_SIG && P.state == PS_WAIT) {
        Sig& S = R.sig[P.wsig & 0x3FFF];
        waiters_unlink(R, S.head, S.tail, p);
    }
    const u32 op = (u32)(R.m.code[P.pc] & 0xFF);
    if (op == OP_RECVT || op == OP_RECVT_X || op == OP_JOIN || op == OP_PWAIT || op == OP_SIG_WAIT) {
        cancel_balker(R, p);
        if ((op == OP_RECVT || op == OP_RECVT_X) && P.phase == 1) select_cleanup(R, p);
        if (exc == EXC_GS_TIMEOUT && op != OP_SIG_WAIT) exc = EXC_ITSIM_TIMEOUT;
    }
    return exc;
}

}  // namespace itsb
