/*
 * itsb_host_common.h -- host-side helpers shared by the CUDA library (itsb_abi.cu) and the test-only host
 * emulation build (tests/hostemu): model validation and the per-replica hot-block layout.
 */
#pragma once
#include <algorithm>
#include <vector>
#include <string.h>
#include <stdio.h>
#include <string>
#include "itsb_core.cuh"

namespace itsb {

static inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

/* Lays the tables of one replica out in a single contiguous block (16-byte aligned regions, total a multiple of
 * 128 bytes so replicas move with aligned 128-bit / bulk copies). */
static inline Layout compute_layout(const itsb_model_desc& d) {
    Layout L;
    memset(&L, 0, sizeof(L));
    const itsb_caps& c = d.caps;
    L.max_ev = c.max_events; L.max_proc = c.max_procs; L.max_sock = c.max_sockets;
    L.n_nodes = d.n_nodes; L.arena_nodes = c.arena_nodes; L.trace_cap = c.trace_records;
    L.log_cap = c.net_event_records;
    uint32_t off = align_up((uint32_t)sizeof(Hdr), 16);
    /* the future-event heap: entry 0 at byte 48 of a 64-byte line, so children 4i + 1 .. 4i + 4 share one line */
    L.off_ev_heap = align_up(off, 64) + 48; off = align_up(L.off_ev_heap + (uint32_t)sizeof(HeapEnt) * L.max_ev, 16);
    L.off_ev_pos = off;    off = align_up(off + 2 * L.max_ev, 16);
    L.off_ev_data = off;   off = align_up(off + 4 * L.max_ev, 16);
    L.off_ev_free = off;   off = align_up(off + 2 * L.max_ev, 16);
    L.off_nq = off;        off = align_up(off + 8 * NQ_CAP, 16);
    L.off_pcb = off;       off = align_up(off + (uint32_t)sizeof(Pcb) * L.max_proc, 16);
    L.off_pcb_free = off;  off = align_up(off + 2 * L.max_proc, 16);
    L.off_sock = off;      off = align_up(off + (uint32_t)sizeof(Sock) * L.max_sock, 16);
    L.off_sock_free = off; off = align_up(off + 2 * L.max_sock, 16);
    L.off_sockpkt = align_up(off, 32); off = L.off_sockpkt + (uint32_t)sizeof(ArenaNode) * L.max_sock;   /* one sector each */
    L.off_sockv = off;     off = align_up(off + (uint32_t)sizeof(SockV) * L.max_sock, 16);
    L.off_node = off;      off = align_up(off + (uint32_t)sizeof(NodeDyn) * L.n_nodes, 16);
    L.off_ndig = align_up(off, 32); off = align_up(L.off_ndig + 4 * L.n_nodes, 16);   /* sector-aligned: 8 nodes per sector */
    L.off_nmask = align_up(off, 32); off = align_up(L.off_nmask + 8 * L.n_nodes, 16);
    L.off_pmask = off;     off = align_up(off + 8 * L.max_proc, 16);
    L.off_stats = off;     off = align_up(off + 8 * ITSB_TM_SIZE, 16);
    /* shipped-itsim tables */
    L.n_ifaces = d.n_ifaces; L.max_sig = c.max_signals; L.max_thr = c.max_threads; L.max_task = c.max_tasks;
    L.off_iface = off;     off = align_up(off + 4 * L.n_ifaces, 16);
    L.off_pcbx = off;      off = align_up(off + (L.max_sig ? (uint32_t)sizeof(PcbX) * L.max_proc : 0), 16);
    L.off_sig = off;       off = align_up(off + (uint32_t)sizeof(Sig) * L.max_sig, 16);
    L.off_sig_free = off;  off = align_up(off + 2 * L.max_sig, 16);
    L.off_thr = off;       off = align_up(off + (uint32_t)sizeof(Thr) * L.max_thr, 16);
    L.off_thr_free = off;  off = align_up(off + 2 * L.max_thr, 16);
    L.off_task = off;      off = align_up(off + (uint32_t)sizeof(Task) * L.max_task, 16);
    L.off_task_free = off; off = align_up(off + 2 * L.max_task, 16);
    L.table_words = c.table_words;
    L.off_tab = off;       off = align_up(off + 4 * L.table_words, 16);
    /* connection records: fill counts of the two correlation windows of every socket */
    L.conn_cap = (c.connection_records + 1u) & ~1u;      /* even: the order keys behind the records stay 8-byte aligned */
    L.conn_window = c.connection_records ? c.connection_window : 0;
    L.address_tags = d.address_tags;
    L.off_sockx = off;     off = align_up(off + (L.conn_cap ? 4 * L.max_sock : 0), 16);
    L.hot_bytes = align_up(off, 128);
    return L;
}

/* bytes of one replica's connection log: conn_cap records followed by conn_cap u32 order keys (itsb_core.cuh conn_order) */
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline size_t conn_log_bytes(uint32_t conn_cap) { return (size_t)conn_cap * (sizeof(itsb_connection) + sizeof(uint32_t)); }

/* Puts the first n records of a replica's log in the order Socket.log_pair was called in: by End_Time, then by the
 * counter value of the event that logged them (records of wake-ups retired in the delivery pass are appended early,
 * under the key their wake-up would have had).  Stable, so records of one event keep their order. */
static inline void order_connections(itsb_connection* recs, const uint32_t* keys, size_t n) {
    bool sorted = true;
    for (size_t i = 1; i < n && sorted; ++i)
        sorted = recs[i - 1].t_end < recs[i].t_end || (recs[i - 1].t_end == recs[i].t_end && keys[i - 1] <= keys[i]);
    if (sorted) return;
    std::vector<uint32_t> idx(n);
    for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
        return recs[a].t_end < recs[b].t_end || (recs[a].t_end == recs[b].t_end && keys[a] < keys[b]);
    });
    std::vector<itsb_connection> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = recs[idx[i]];
    memcpy(recs, out.data(), n * sizeof(itsb_connection));
}

static inline bool validate_model(const itsb_model_desc& d, const void* const* blobs, const size_t* sizes,
                                  std::string& why) {
    char buf[256];
    if (d.abi_version != ITSB_ABI_VERSION) { why = "abi_version mismatch"; return false; }
    const size_t want[ITSB_NUM_BLOBS] = {
        (size_t)d.n_code * 8, (size_t)d.n_entries * sizeof(itsb_entry), (size_t)d.n_dists * sizeof(itsb_dist),
        (size_t)d.n_consts * 8, (size_t)d.n_nets * sizeof(itsb_net), (size_t)d.n_nodes * sizeof(itsb_node),
        (size_t)d.n_init * sizeof(itsb_initproc), (size_t)d.n_filters * sizeof(itsb_filter),
        (size_t)d.n_ifaces * sizeof(itsb_iface), (size_t)d.n_routes * sizeof(itsb_route), (size_t)d.n_members * 4};
    for (int i = 0; i < ITSB_NUM_BLOBS; ++i) {
        if (sizes[i] != want[i] || (want[i] && !blobs[i])) {
            snprintf(buf, sizeof buf, "blob %d: %zu bytes given, %zu expected", i, sizes[i], want[i]);
            why = buf; return false;
        }
    }
    const itsb_caps& c = d.caps;
    if (c.max_events < 8 || c.max_events > 65535 || c.max_procs < 1 || c.max_procs > 8192 || c.max_sockets < 1 ||
        c.max_sockets > 0x3FFF || c.arena_nodes < 1 || c.arena_nodes >= (1u << 28)) {
        why = "capacity out of range"; return false;
    }
    if (c.connection_records && (c.connection_window < 1 || c.connection_window > 4096)) { why = "connection_window must be 1..4096"; return false; }
    if (d.n_nodes == 0 || d.n_nodes > 0x3FFF || d.n_code == 0 || d.n_code > 65535) { why = "table size out of range"; return false; }
    const uint64_t* code = (const uint64_t*)blobs[ITSB_BLOB_CODE];
    for (uint32_t i = 0; i < d.n_code; ++i) {
        uint32_t op = (uint32_t)(code[i] & 0xFF), imm = (uint32_t)(code[i] >> 32);
        if (op >= OP_COUNT) { snprintf(buf, sizeof buf, "code[%u]: bad opcode %u", i, op); why = buf; return false; }
        bool has_dist = op == OP_DRAWI || op == OP_DRAW_ADDI;
        if (has_dist && imm >= d.n_dists) { snprintf(buf, sizeof buf, "code[%u]: dist out of range", i); why = buf; return false; }
        if (op == OP_ADVANCE && (imm & 0xFFFF) >= d.n_dists) { why = "ADVANCE dist out of range"; return false; }
        if (op == OP_ADVANCE_C && (imm & 0xFFFF) >= d.n_consts) { why = "ADVANCE_C const out of range"; return false; }
        if (op == OP_SPAWN && imm >= d.n_entries) { why = "SPAWN entry out of range"; return false; }
        if (c.connection_records && (op == OP_RECV || op == OP_RECV_FILTER || op == OP_SEND || op == OP_FORWARD || op == OP_CLOSE)) {
            why = "connection records need programs assembled with the logged socket instructions (OP_*_LOG)"; return false;
        }
        if ((op == OP_STAT || op == OP_STAT_ADD || op == OP_STAT_FIRST) && imm >= ITSB_TM_USER_SLOTS) {
            why = "model counter out of range"; return false;
        }
        if (op == OP_RECV_FILTER || op == OP_RECV_FILTER_LOG) {
            uint32_t a = (uint32_t)((code[i] >> 8) & 0xFF), b = (uint32_t)((code[i] >> 16) & 0xFF);
            if (a > 3 || b >= d.n_filters || (imm & 0xFFFF) >= d.n_code) { why = "RECV_FILTER operand out of range"; return false; }
        }
        if ((op == OP_JMP || op == OP_JEQ || op == OP_JNE || op == OP_JASLEEP || op == OP_LIST_TAKE) && imm >= d.n_code) {
            snprintf(buf, sizeof buf, "code[%u]: jump out of range", i); why = buf; return false;
        }
        if ((op == OP_JEQI || op == OP_JNEI) && (imm & 0xFFFF) >= d.n_code) { why = "jump out of range"; return false; }
    }
    /* generator parameters the device trusts: a CHOICE over nothing never terminates its rejection loop, an
       exponential with a non-positive or non-finite rate divides by zero */
    const itsb_dist* ds = (const itsb_dist*)blobs[ITSB_BLOB_DISTS];
    for (uint32_t i = 0; i < d.n_dists; ++i) {
        const itsb_dist& D = ds[i];
        const double ps[6] = {D.p0, D.p1, D.slope, D.shift, D.lower, D.upper};
        for (double v : ps) if (!(v == v) || v > 1.7e308 || v < -1.7e308) { snprintf(buf, sizeof buf, "dist[%u]: non-finite parameter", i); why = buf; return false; }
        if (D.kind > ITSB_DIST_CHOICE) { snprintf(buf, sizeof buf, "dist[%u]: unknown kind", i); why = buf; return false; }
        if (D.kind == ITSB_DIST_CHOICE && !(D.p0 >= 1.0 && D.p0 <= 2147483648.0)) { snprintf(buf, sizeof buf, "dist[%u]: CHOICE over %g values", i, D.p0); why = buf; return false; }
        if (D.kind == ITSB_DIST_EXPO && !(D.p0 > 0.0)) { snprintf(buf, sizeof buf, "dist[%u]: exponential rate must be positive", i); why = buf; return false; }
    }
    const itsb_entry* en = (const itsb_entry*)blobs[ITSB_BLOB_ENTRIES];
    for (uint32_t i = 0; i < d.n_entries; ++i) if (en[i].pc >= d.n_code) { why = "entry pc out of range"; return false; }
    const itsb_net* nets = (const itsb_net*)blobs[ITSB_BLOB_NETS];
    for (uint32_t i = 0; i < d.n_nets; ++i) {
        if (nets[i].lat_dist >= d.n_dists || nets[i].bw_dist >= d.n_dists || nets[i].mode > ITSB_NET_LINK) {
            why = "net table out of range"; return false;
        }
    }
    const itsb_node* nodes = (const itsb_node*)blobs[ITSB_BLOB_NODES];
    for (uint32_t i = 0; i < d.n_nodes; ++i) if (nodes[i].net >= d.n_nets) { why = "node.net out of range"; return false; }
    if (d.n_nets == 0) { why = "no network"; return false; }
    const itsb_initproc* ip = (const itsb_initproc*)blobs[ITSB_BLOB_INIT];
    for (uint32_t i = 0; i < d.n_init; ++i) if (ip[i].entry >= d.n_entries || ip[i].node >= d.n_nodes) { why = "init proc out of range"; return false; }
    if (d.n_init > c.max_procs) { why = "more installed processes than slots"; return false; }
    /* shipped-itsim tables */
    const bool machine = d.n_ifaces != 0;
    if (machine && (c.max_signals == 0 || c.max_threads == 0 || c.max_tasks == 0)) { why = "shipped-itsim model without thread/task/signal tables"; return false; }
    if (c.max_signals > 0x3FFF || c.max_threads > 0xFFFE || c.max_tasks > 0xFFFE) { why = "table capacity out of range"; return false; }
    const itsb_iface* ifs = (const itsb_iface*)blobs[ITSB_BLOB_IFACES];
    for (uint32_t i = 0; i < d.n_ifaces; ++i)
        if (ifs[i].node >= d.n_nodes || ifs[i].net >= d.n_nets || ifs[i].first_route + ifs[i].n_routes > d.n_routes) { why = "interface table out of range"; return false; }
    const uint32_t* mem = (const uint32_t*)blobs[ITSB_BLOB_MEMBERS];
    for (uint32_t i = 0; i < d.n_members; ++i) if (mem[i] >= d.n_nodes) { why = "member out of range"; return false; }
    for (uint32_t i = 0; i < d.n_nets; ++i) {
        if (nets[i].mode == ITSB_NET_LINK && nets[i].first_node + nets[i].n_nodes > d.n_members) { why = "link members out of range"; return false; }
        if (nets[i].mode == ITSB_NET_OVERRIDE && nets[i].first_node + nets[i].n_nodes > d.n_nodes) { why = "net table out of range"; return false; }
    }
    for (uint32_t i = 0; i < d.n_nodes; ++i) {
        uint32_t fi = nodes[i].flags & 0xFFFF, ni = nodes[i].flags >> 16;
        if (fi + ni > d.n_ifaces) { why = "node interfaces out of range"; return false; }
    }
    for (uint32_t i = 0; i < d.n_init; ++i)
        if (ip[i].mode == ITSB_INIT_RUN_PROC && (!machine || ip[i].delay_const >= d.n_consts)) { why = "run_proc init entry invalid"; return false; }
    for (uint32_t i = 0; i < d.n_init; ++i) {
        if (ip[i].mode > ITSB_INIT_ADD_SAME) { why = "init mode out of range"; return false; }
        if (ip[i].mode == ITSB_INIT_ADD_DRAWN && ip[i].delay_const >= d.n_dists) { why = "drawn init entry without a generator"; return false; }
        if (ip[i].mode == ITSB_INIT_ADD_DRAWN) {     /* node + draw indexes the node tables: the candidates are node .. node + n - 1 */
            const itsb_dist& D = ds[ip[i].delay_const];
            if (D.kind != ITSB_DIST_CHOICE || D.flags != 0 || (double)ip[i].node + D.p0 > (double)d.n_nodes) {
                why = "drawn init entry: the generator must be a plain CHOICE over consecutive nodes inside the node table"; return false;
            }
        }
        if (ip[i].mode == ITSB_INIT_ADD_SAME && i == 0) { why = "ITSB_INIT_ADD_SAME without a drawn entry before it"; return false; }
    }
    bool has_balk = false, has_wait_one = false;
    for (uint32_t i = 0; i < d.n_entries; ++i) { has_balk |= en[i].prog == PROG_HIDDEN_BALK; has_wait_one |= en[i].prog == PROG_HIDDEN_WAIT_ONE; }
    if (machine && !(has_balk && has_wait_one)) { why = "shipped-itsim model without the greensim helper programs"; return false; }
    return true;
}

}  // namespace itsb
