/*
 * TEST INFRASTRUCTURE ONLY -- host emulation of the device interpreter.
 *
 * Compiles itsim_b200/csrc/itsb_core.cuh with ITSB_W == 1 (one "lane") as plain C++ so the unit tests that run
 * without a GPU (pytest -m "not gpu") can check the interpreter's semantics against the oracle's golden traces.
 * It exports the same itsb_* symbols as the product library so the same ctypes wrapper drives it, but it is built
 * into tests/hostemu/ and only tests load it (explicitly, by path).  The product package never does: without the
 * CUDA library it fails loudly (itsim_b200/_lib.py).
 */
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../itsim_b200/csrc/itsb_host_common.h"

using namespace itsb;

struct itsb_engine {
    itsb_model_desc desc;
    std::vector<uint64_t> code;
    std::vector<itsb_entry> entries;
    std::vector<itsb_dist> dists;
    std::vector<double> consts;
    std::vector<itsb_net> nets;
    std::vector<itsb_node> nodes;
    std::vector<itsb_initproc> init;
    std::vector<itsb_filter> filters;
    std::vector<itsb_iface> ifaces;
    std::vector<itsb_route> routes;
    std::vector<uint32_t> members;
    Model model;
    Layout layout;
    uint64_t n_replicas = 0, seed = 0, first = 0;
    std::vector<unsigned char> hot;
    std::vector<ArenaNode> arena;
    std::vector<uint32_t> arena_free;
    std::vector<itsb_trace_rec> trace;
    std::vector<itsb_net_event> netlog;
    std::vector<unsigned char> connlog;      /* per replica: conn_cap records, then conn_cap order keys */
    std::vector<ConnEnt> conntab;
    uint64_t last_events = 0, launches = 0;
    /* itsb_drain_*: two alternating sets, like the device library */
    struct DrainSet { std::vector<itsb_net_event> net; std::vector<itsb_connection> conn; std::vector<uint64_t> first;
                      uint64_t dropped[2] = {0, 0}; bool pending = false; } drain[2];
    unsigned drain_next = 0;
    std::string err;
    bool loaded = false;
};

extern "C" {

int itsb_abi_version(void) { return ITSB_ABI_VERSION; }

int itsb_create(int, itsb_engine** out) { *out = new itsb_engine(); return ITSB_OK; }

int itsb_load_model(itsb_engine* e, const itsb_model_desc* d, const void* const* blobs, const size_t* sizes) {
    if (!validate_model(*d, blobs, sizes, e->err)) return ITSB_ERR_BAD_MODEL;
    e->desc = *d;
    e->code.assign((const uint64_t*)blobs[0], (const uint64_t*)blobs[0] + d->n_code);
    e->entries.assign((const itsb_entry*)blobs[1], (const itsb_entry*)blobs[1] + d->n_entries);
    e->dists.assign((const itsb_dist*)blobs[2], (const itsb_dist*)blobs[2] + d->n_dists);
    e->consts.assign((const double*)blobs[3], (const double*)blobs[3] + d->n_consts);
    e->nets.assign((const itsb_net*)blobs[4], (const itsb_net*)blobs[4] + d->n_nets);
    e->nodes.assign((const itsb_node*)blobs[5], (const itsb_node*)blobs[5] + d->n_nodes);
    e->init.assign((const itsb_initproc*)blobs[6], (const itsb_initproc*)blobs[6] + d->n_init);
    e->filters.assign((const itsb_filter*)blobs[7], (const itsb_filter*)blobs[7] + d->n_filters);
    e->ifaces.assign((const itsb_iface*)blobs[8], (const itsb_iface*)blobs[8] + d->n_ifaces);
    e->routes.assign((const itsb_route*)blobs[9], (const itsb_route*)blobs[9] + d->n_routes);
    e->members.assign((const uint32_t*)blobs[10], (const uint32_t*)blobs[10] + d->n_members);
    Model& m = e->model;
    m.code = e->code.data(); m.entries = e->entries.data(); m.dists = e->dists.data(); m.consts = e->consts.data();
    m.nets = e->nets.data(); m.nodes = e->nodes.data(); m.init = e->init.data(); m.filters = e->filters.data();
    m.ifaces = e->ifaces.data(); m.routes = e->routes.data(); m.members = e->members.data();
    m.n_ifaces = d->n_ifaces; m.n_routes = d->n_routes; m.n_members = d->n_members;
    m.n_code = d->n_code; m.n_entries = d->n_entries; m.n_dists = d->n_dists; m.n_consts = d->n_consts;
    m.n_nets = d->n_nets; m.n_nodes = d->n_nodes; m.n_init = d->n_init; m.n_filters = d->n_filters;
    e->layout = compute_layout(*d);
    e->loaded = true;
    return ITSB_OK;
}

static void bind(itsb_engine* e, Rep& R, uint64_t r) {
    const Layout& L = e->layout;
    Binding B;
    B.arena = e->arena.data() + r * L.arena_nodes;
    B.arena_free = e->arena_free.data() + r * L.arena_nodes;
    B.trace = L.trace_cap ? e->trace.data() + r * L.trace_cap : nullptr;
    B.netlog = L.log_cap ? e->netlog.data() + r * L.log_cap : nullptr;
    B.connlog = L.conn_cap ? (itsb_connection*)(e->connlog.data() + r * conn_log_bytes(L.conn_cap)) : nullptr;
    B.conntab = L.conn_cap ? e->conntab.data() + r * ((size_t)L.max_sock * 2 * L.conn_window) : nullptr;
    bind_views(R, e->hot.data() + r * L.hot_bytes, &L, &e->model, B, e->seed);
}

int itsb_init_replicas(itsb_engine* e, uint64_t n, uint64_t seed, uint64_t first) {
    if (!e->loaded) return ITSB_ERR_STATE;
    e->n_replicas = n; e->seed = seed; e->first = first;
    e->hot.assign(n * e->layout.hot_bytes, 0);
    e->arena.assign(n * e->layout.arena_nodes, ArenaNode());
    e->arena_free.resize(n * e->layout.arena_nodes);
    e->trace.assign(n * e->layout.trace_cap, itsb_trace_rec());
    e->netlog.assign(n * e->layout.log_cap, itsb_net_event());
    e->connlog.assign(n * conn_log_bytes(e->layout.conn_cap), 0);
    e->conntab.assign(n * ((size_t)e->layout.max_sock * 2 * e->layout.conn_window), ConnEnt());
    for (uint64_t r = 0; r < n; ++r) {
        Rep R; bind(e, R, r);
        for (uint32_t i = 0; i < e->layout.arena_nodes; ++i) R.arena_free[i] = e->layout.arena_nodes - 1 - i;
        init_replica(R, first + r);
    }
    return ITSB_OK;
}

int itsb_run_async(itsb_engine* e, double duration) {
    if (!e->n_replicas) return ITSB_ERR_STATE;
    for (uint64_t r = 0; r < e->n_replicas; ++r) {
        Rep R; bind(e, R, r);
        e->last_events += run_replica(R, duration);
    }
    e->launches += 1;
    return ITSB_OK;
}

int64_t itsb_first_failed_replica(itsb_engine* e) {
    for (uint64_t r = 0; r < e->n_replicas; ++r)
        if (((Hdr*)(e->hot.data() + r * e->layout.hot_bytes))->status != 0) return (int64_t)r;
    return -1;
}

int itsb_sync(itsb_engine* e, uint64_t* events_done) {
    if (events_done) *events_done = e->last_events;
    e->last_events = 0;
    int64_t bad = itsb_first_failed_replica(e);
    if (bad >= 0) return ((Hdr*)(e->hot.data() + bad * e->layout.hot_bytes))->status;
    return ITSB_OK;
}

int itsb_run(itsb_engine* e, double duration, uint64_t* events_done) {
    int rc = itsb_run_async(e, duration);
    if (rc) return rc;
    return itsb_sync(e, events_done);
}

int itsb_last_run_ms(itsb_engine*, float* ms) { *ms = 0.f; return ITSB_OK; }
int itsb_launch_count(itsb_engine* e, uint64_t* n) { *n = e->launches; return ITSB_OK; }
const char* itsb_kernel_name(itsb_engine*) { return "host emulation (tests only)"; }

int itsb_read_now(itsb_engine* e, uint64_t first, uint64_t n, double* dst) {
    if (first + n > e->n_replicas) return ITSB_ERR_STATE;
    for (uint64_t i = 0; i < n; ++i) dst[i] = ((Hdr*)(e->hot.data() + (first + i) * e->layout.hot_bytes))->now;
    return ITSB_OK;
}

int itsb_read_trace(itsb_engine* e, uint64_t replica, itsb_trace_rec* dst, size_t cap, size_t* n_total) {
    if (replica >= e->n_replicas) return ITSB_ERR_STATE;
    Hdr* H = (Hdr*)(e->hot.data() + replica * e->layout.hot_bytes);
    if (n_total) *n_total = H->trace_count;
    size_t n = H->trace_count < e->layout.trace_cap ? H->trace_count : e->layout.trace_cap;
    if (n > cap) n = cap;
    if (n) memcpy(dst, e->trace.data() + replica * e->layout.trace_cap, n * sizeof(itsb_trace_rec));
    return ITSB_OK;
}

int itsb_read_net_events(itsb_engine* e, uint64_t replica, itsb_net_event* dst, size_t cap, size_t* n_total) {
    if (replica >= e->n_replicas) return ITSB_ERR_STATE;
    Hdr* H = (Hdr*)(e->hot.data() + replica * e->layout.hot_bytes);
    if (n_total) *n_total = H->log_count;
    size_t n = H->log_count < e->layout.log_cap ? H->log_count : e->layout.log_cap;
    if (n > cap) n = cap;
    if (n) memcpy(dst, e->netlog.data() + replica * e->layout.log_cap, n * sizeof(itsb_net_event));
    return ITSB_OK;
}

int itsb_histogram(itsb_engine* e, uint32_t slot, uint64_t lo, uint64_t width, uint32_t nbins, int only_nonzero,
                   void* comm, uint64_t* dst) {
    if (!e->n_replicas || !dst || slot >= ITSB_TM_SIZE || width == 0 || nbins == 0 || nbins > 4096 || comm) return ITSB_ERR_STATE;
    for (uint32_t b = 0; b < nbins; ++b) dst[b] = 0;
    for (uint64_t r = 0; r < e->n_replicas; ++r) {
        const uint64_t v = ((const uint64_t*)(e->hot.data() + r * e->layout.hot_bytes + e->layout.off_stats))[slot];
        if (only_nonzero && v == 0) continue;
        uint64_t b = v <= lo ? 0 : (v - lo) / width;
        dst[b >= nbins ? nbins - 1 : b] += 1;
    }
    return ITSB_OK;
}

int itsb_read_connections(itsb_engine* e, uint64_t replica, itsb_connection* dst, size_t cap, size_t* n_total) {
    if (replica >= e->n_replicas) return ITSB_ERR_STATE;
    Hdr* H = (Hdr*)(e->hot.data() + replica * e->layout.hot_bytes);
    if (n_total) *n_total = H->conn_count;
    size_t n = H->conn_count < e->layout.conn_cap ? H->conn_count : e->layout.conn_cap;
    if (n > cap) n = cap;
    if (n) {
        const size_t held = H->conn_count < e->layout.conn_cap ? H->conn_count : e->layout.conn_cap;
        const unsigned char* base = e->connlog.data() + replica * conn_log_bytes(e->layout.conn_cap);
        std::vector<itsb_connection> recs(held);
        std::vector<uint32_t> keys(held);
        memcpy(recs.data(), base, held * sizeof(itsb_connection));
        memcpy(keys.data(), base + (size_t)e->layout.conn_cap * sizeof(itsb_connection), held * sizeof(uint32_t));
        order_connections(recs.data(), keys.data(), held);
        memcpy(dst, recs.data(), n * sizeof(itsb_connection));
    }
    return ITSB_OK;
}

int itsb_drain_begin(itsb_engine* e) {
    if (!e->n_replicas) return ITSB_ERR_STATE;
    const Layout& L = e->layout;
    if (L.log_cap == 0 && L.conn_cap == 0) return ITSB_ERR_STATE;
    auto& D = e->drain[e->drain_next & 1];
    if (D.pending) return ITSB_ERR_STATE;
    const uint64_t n = e->n_replicas;
    D.net.clear(); D.conn.clear(); D.first.assign(2 * (n + 1), 0); D.dropped[0] = D.dropped[1] = 0;
    for (uint64_t r = 0; r < n; ++r) {
        Hdr* H = (Hdr*)(e->hot.data() + r * L.hot_bytes);
        const uint32_t a = H->log_count < L.log_cap ? H->log_count : L.log_cap;
        const uint32_t b = H->conn_count < L.conn_cap ? H->conn_count : L.conn_cap;
        D.dropped[0] += H->log_count - a; D.dropped[1] += H->conn_count - b;
        D.first[r] = D.net.size(); D.first[n + 1 + r] = D.conn.size();
        if (a) D.net.insert(D.net.end(), e->netlog.begin() + r * L.log_cap, e->netlog.begin() + r * L.log_cap + a);
        if (b) {
            const unsigned char* base = e->connlog.data() + r * conn_log_bytes(L.conn_cap);
            std::vector<itsb_connection> recs(b);
            std::vector<uint32_t> keys(b);
            memcpy(recs.data(), base, b * sizeof(itsb_connection));
            memcpy(keys.data(), base + (size_t)L.conn_cap * sizeof(itsb_connection), b * sizeof(uint32_t));
            order_connections(recs.data(), keys.data(), b);
            D.conn.insert(D.conn.end(), recs.begin(), recs.end());
        }
        H->log_count = 0; H->conn_count = 0;
    }
    D.first[n] = D.net.size(); D.first[2 * n + 1] = D.conn.size();
    D.pending = true;
    e->drain_next += 1;
    e->launches += 3;
    return ITSB_OK;
}

int itsb_drain_end(itsb_engine* e, itsb_drain* out) {
    itsb_engine::DrainSet* D = nullptr;
    for (unsigned k = 0; k < 2; ++k) { auto& C = e->drain[(e->drain_next + k) & 1]; if (C.pending) { D = &C; break; } }
    if (!D) return ITSB_ERR_STATE;
    const uint64_t n = e->n_replicas;
    out->net = D->net.data(); out->n_net = D->net.size(); out->net_first = D->first.data();
    out->conn = D->conn.data(); out->n_conn = D->conn.size(); out->conn_first = D->first.data() + (n + 1);
    out->net_dropped = D->dropped[0]; out->conn_dropped = D->dropped[1];
    out->d2h_bytes = out->n_net * sizeof(itsb_net_event) + out->n_conn * sizeof(itsb_connection) + 2 * (n + 1) * 8 + 32;
    out->pack_ms = 0.0;
    D->pending = false;
    return ITSB_OK;
}

int itsb_read_replica_telemetry(itsb_engine* e, uint64_t replica, uint64_t* dst, size_t n) {
    if (replica >= e->n_replicas || n < ITSB_TM_SIZE) return ITSB_ERR_STATE;
    memcpy(dst, e->hot.data() + replica * e->layout.hot_bytes + e->layout.off_stats, 8 * ITSB_TM_SIZE);
    return ITSB_OK;
}

int itsb_read_telemetry(itsb_engine* e, uint64_t* dst, size_t n) {
    if (n < ITSB_TM_SIZE) return ITSB_ERR_STATE;
    for (size_t i = 0; i < ITSB_TM_SIZE; ++i) dst[i] = 0;
    for (uint64_t r = 0; r < e->n_replicas; ++r) {
        const uint64_t* s = (const uint64_t*)(e->hot.data() + r * e->layout.hot_bytes + e->layout.off_stats);
        for (size_t i = 0; i < ITSB_TM_SIZE; ++i) dst[i] += s[i];
    }
    return ITSB_OK;
}

int itsb_allreduce_telemetry(itsb_engine* e, void*, uint64_t* dst, size_t n) { return itsb_read_telemetry(e, dst, n); }
int itsb_allreduce_u64(itsb_engine*, void*, uint64_t*, size_t, int) { return ITSB_ERR_NCCL; }
int itsb_nccl_unique_id(void*, size_t) { return ITSB_ERR_NCCL; }
int itsb_nccl_comm_init(itsb_engine*, int, int, const void*, void**) { return ITSB_ERR_NCCL; }
int itsb_nccl_comm_destroy(void*) { return ITSB_ERR_NCCL; }

int itsb_replica_status(itsb_engine* e, uint64_t replica, int32_t* status, uint64_t* detail) {
    if (replica >= e->n_replicas) return ITSB_ERR_STATE;
    Hdr* H = (Hdr*)(e->hot.data() + replica * e->layout.hot_bytes);
    if (status) *status = H->status;
    if (detail) *detail = H->detail;
    return ITSB_OK;
}

int itsb_state_bytes(itsb_engine* e, uint64_t* hot, uint64_t* arena) {
    if (hot) *hot = e->layout.hot_bytes;
    if (arena) *arena = (uint64_t)e->layout.arena_nodes * (sizeof(ArenaNode) + 4);
    return ITSB_OK;
}

int itsb_eval_log(itsb_engine*, const double* x, double* y, size_t n) {
    for (size_t i = 0; i < n; ++i) y[i] = log_f64(x[i]);
    return ITSB_OK;
}

const char* itsb_last_error(itsb_engine* e) { return e ? e->err.c_str() : ""; }
void itsb_destroy(itsb_engine* e) { delete e; }

}  // extern "C"
