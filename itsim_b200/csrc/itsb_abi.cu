/*
 * itsb_abi.cu -- libitsim_b200.so: the C ABI of include/itsb.h and the sm_100a kernels behind it.
 *
 * Kernels (all hand-written for B200; no tensor cores -- nothing here is a contraction):
 *   k_run          persistent, one CTA per SM, one warp per resident replica.  A warp claims a replica from a
 *                  global ticket, pulls its hot block HBM -> shared memory with one TMA bulk copy
 *                  (cp.async.bulk + mbarrier), runs Simulator.run(duration) on it (itsb_core.cuh), pushes the block
 *                  back with a bulk store, claims the next.  Static model tables are staged once per CTA.
 *                  A model whose replica does not fit one warp's share of shared memory (the 1 040-node LAN of
 *                  config C4: ~650 KB per replica) is advanced in place instead: the same warp-per-replica code
 *                  on the hot block in HBM, cached by L1/L2 (each replica is touched by one warp of one SM only).
 *   k_make_template / k_fan_out   build the t = 0 state once, replicate it to every replica with 128-bit copies.
 *   k_reduce_telemetry            column sums of the per-replica counters / histograms.
 *
 * Layout in HBM: replica r owns hot[r * hot_bytes] (event pool, process table, sockets, nodes, in-flight packets,
 * counters: everything the warp touches per event), arena[r * arena_nodes] (32-byte queue nodes, one DRAM sector
 * each) with its free stack, and optionally trace[r * trace_cap].
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "itsb_kernel.cuh"

extern "C" __global__ void k_run_in_place(const RunParams p);   /* itsb_inplace.cu */
extern "C" __global__ void k_run_wide(const RunParams p);       /* itsb_wide.cu: 13..16 warps per CTA */
extern "C" __global__ void k_run_vote(const RunParams p);       /* itsb_lanes.cu: a replica per lane + per-warp majority key */
extern "C" __global__ void k_run_vote_hbm(const RunParams p);   /* itsb_lanes_hbm.cu: the same for models whose tables stay in HBM */
#if defined(ITSB_EXPERIMENTS)
/* Variants measured slower in round 1 (profiles/README.md holds the tables); built only with -DITSB_EXPERIMENTS
 * (python __graft_entry__.py --experiments), never part of the product library. */
extern "C" __global__ void k_run_phased(const RunParams p);     /* itsb_phased.cu: warps aligned on the loop's two halves */
extern "C" __global__ void k_run_lanes(const RunParams p);      /* itsb_lanes.cu: a replica per lane, no vote */
extern "C" __global__ void k_run_sorted(const RunParams p);     /* itsb_sorted.cu: lanes regrouped by next event, CTA-wide */
#endif
/* itsb_pool.cu: k_run_pool<threads, CTAs per SM, slots, ring> -- a replica per lane, replicas change warps through per-kind
 * rings in shared memory; the shapes are instantiated there */
extern "C" int itsb_pool_shapes(void);
extern "C" const char* itsb_pool_shape_name(int shape);
extern "C" bool itsb_pool_prepare(int shape, size_t table_bytes);
extern "C" cudaError_t itsb_pool_launch(int shape, int n_sms, size_t table_bytes, RunParams rp, cudaStream_t stream);
extern "C" bool itsb_pool_prepare_hbm(int shape, size_t table_bytes);          /* itsb_pool_hbm.cu: tables left in HBM */
extern "C" cudaError_t itsb_pool_launch_hbm(int shape, int n_sms, size_t table_bytes, RunParams rp, cudaStream_t stream);

/* up to 12 warps per CTA (what shared memory allows the flagship model): 168 registers per thread, no spills */
extern "C" __global__ void __launch_bounds__(384, 1) k_run(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    run_warps<false>(p, smem);
}

/* Builds the t = 0 hot block once (replica id patched in by k_fan_out). */
extern "C" __global__ void k_make_template(const RunParams p, const Model* model, unsigned char* tmpl,
                                           ArenaNode* arena0, uint32_t* free0) {
    Rep R;
    Binding B;
    B.arena = arena0; B.arena_free = free0; B.trace = nullptr; B.netlog = nullptr; B.connlog = nullptr; B.conntab = nullptr;
    bind_views(R, tmpl, &p.layout, model, B, p.seed);
    init_replica(R, 0);
}

/* Replicates the template into every replica's hot block, fills the arena free stacks and copies the arena nodes
 * the template already uses (installed processes beyond the now-queue ring start life in the spill chain; the free
 * stack hands out node 0, 1, 2, ... so the used nodes are the prefix [0, used)). */
extern "C" __global__ void k_fan_out(const RunParams p, const unsigned char* tmpl, const ArenaNode* tmpl_arena,
                                     uint64_t first_replica_id) {
    const uint32_t vec = p.layout.hot_bytes / 16;
    const uint4* src = (const uint4*)tmpl;
    const uint32_t used = p.layout.arena_nodes - ((const Hdr*)tmpl)->arena_free_top;
    for (uint64_t r = blockIdx.x; r < p.n_replicas; r += gridDim.x) {
        uint4* dst = (uint4*)(p.hot + r * p.layout.hot_bytes);
        for (uint32_t i = threadIdx.x; i < vec; i += blockDim.x) dst[i] = src[i];
        uint4* adst = (uint4*)(p.arena + r * p.layout.arena_nodes);
        for (uint32_t i = threadIdx.x; i < used * 2; i += blockDim.x) adst[i] = ((const uint4*)tmpl_arena)[i];
        uint32_t* fr = p.arena_free + r * p.layout.arena_nodes;
        for (uint32_t i = threadIdx.x; i < p.layout.arena_nodes; i += blockDim.x) fr[i] = p.layout.arena_nodes - 1 - i;
        __syncthreads();
        if (threadIdx.x == 0) {
            Hdr* H = (Hdr*)dst;
            uint64_t id = first_replica_id + r;
            H->replica_lo = (uint32_t)id;
            H->replica_hi = (uint32_t)(id >> 32);
        }
    }
}

extern "C" __global__ void k_reduce_telemetry(const unsigned char* hot, uint32_t hot_bytes, uint32_t off_stats,
                                              uint64_t n, unsigned long long* out) {
    const uint32_t c = threadIdx.x;
    if (c >= ITSB_TM_SIZE) return;
    unsigned long long acc = 0;
    for (uint64_t r = blockIdx.x; r < n; r += gridDim.x)
        acc += ((const unsigned long long*)(hot + r * hot_bytes + off_stats))[c];
    if (acc) atomicAdd(out + c, acc);
}

/* the engine's logarithm on an array (itsb_eval_log): what the variates call, reachable from a parity test */
extern "C" __global__ void k_eval_log(double* v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = log_f64(v[i]);
}

/* ---- streaming the record logs home (itsb_drain_begin / itsb_drain_end) ---------------------------------------------- */
/* records each replica holds since the last drain (capped at its capacity) and what it produced beyond that */
extern "C" __global__ void k_log_count(const unsigned char* hot, uint32_t hot_bytes, uint64_t n, uint32_t log_cap,
                                       uint32_t conn_cap, unsigned long long* n_net, unsigned long long* n_conn,
                                       unsigned long long* dropped) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const Hdr* H = (const Hdr*)(hot + r * hot_bytes);
    const uint32_t a = H->log_count, b = H->conn_count;
    n_net[r] = a < log_cap ? a : log_cap;
    n_conn[r] = b < conn_cap ? b : conn_cap;
    if (a > log_cap) atomicAdd(dropped, (unsigned long long)(a - log_cap));
    if (b > conn_cap) atomicAdd(dropped + 1, (unsigned long long)(b - conn_cap));
}

/* exclusive prefix sums of the two count arrays, in place, totals behind them ([n]); one CTA: the arrays are small */
extern "C" __global__ void __launch_bounds__(1024) k_log_scan(unsigned long long* a, unsigned long long* b, uint64_t n) {
    __shared__ unsigned long long warp_sums[2][32];
    __shared__ unsigned long long carry[2];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        unsigned long long v[2] = {i < n ? a[i] : 0ull, i < n ? b[i] : 0ull}, x[2];
        for (int k = 0; k < 2; ++k) {
            x[k] = v[k];
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x[k], d);
                if (lane >= (uint32_t)d) x[k] += y;
            }
            if (lane == 31) warp_sums[k][warp] = x[k];
        }
        __syncthreads();
        if (warp == 0) {
            for (int k = 0; k < 2; ++k) {
                unsigned long long w = warp_sums[k][lane];
                for (int d = 1; d < 32; d <<= 1) {
                    unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, w, d);
                    if (lane >= (uint32_t)d) w += y;
                }
                warp_sums[k][lane] = w;
            }
        }
        __syncthreads();
        for (int k = 0; k < 2; ++k) {
            const unsigned long long before = carry[k] + (warp ? warp_sums[k][warp - 1] : 0ull) + x[k] - v[k];
            if (i < n) (k ? b : a)[i] = before;
        }
        __syncthreads();
        if (threadIdx.x == 0) { carry[0] += warp_sums[0][31]; carry[1] += warp_sums[1][31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { a[n] = carry[0]; b[n] = carry[1]; }
}

/* One CTA per replica: its records go to their place in the packed staging buffers, then its log starts over.
 * Connection records are appended in processing order except those of wake-ups retired in the delivery pass, which
 * are logged early under the counter value their wake-up would have had; log_pair's call order is (End_Time, that
 * value).  End_Time never decreases along the log, so only records that share an End_Time can be out of place: each
 * record finds its rank inside its run of equal End_Times (the runs are the events of one instant: short). */
extern "C" __global__ void __launch_bounds__(128) k_log_gather(
    unsigned char* hot, uint32_t hot_bytes, uint64_t n, const itsb_net_event* netlog, uint32_t log_cap,
    const unsigned char* connlog, uint32_t conn_cap, const unsigned long long* net_first,
    const unsigned long long* conn_first, itsb_net_event* out_net, itsb_connection* out_conn) {
    const uint64_t r = blockIdx.x;
    if (r >= n) return;
    Hdr* H = (Hdr*)(hot + r * hot_bytes);
    const uint32_t n_net = (uint32_t)(net_first[r + 1] - net_first[r]);
    const uint32_t n_conn = (uint32_t)(conn_first[r + 1] - conn_first[r]);
    if (n_net) {
        const uint4* src = (const uint4*)(netlog + r * log_cap);
        uint4* dst = (uint4*)(out_net + net_first[r]);
        for (uint32_t i = threadIdx.x; i < n_net * 2; i += blockDim.x) dst[i] = src[i];
    }
    if (n_conn) {
        const unsigned char* base = connlog + r * conn_log_bytes(conn_cap);
        const itsb_connection* recs = (const itsb_connection*)base;
        const uint32_t* keys = (const uint32_t*)(base + (size_t)conn_cap * sizeof(itsb_connection));
        itsb_connection* dst = out_conn + conn_first[r];
        for (uint32_t i = threadIdx.x; i < n_conn; i += blockDim.x) {
            const itsb_connection C = recs[i];
            const uint32_t key = keys[i];
            uint32_t lo = i, rank = 0;
            while (lo > 0 && recs[lo - 1].t_end == C.t_end) { lo -= 1; if (keys[lo] <= key) rank += 1; }
            for (uint32_t j = i + 1; j < n_conn && recs[j].t_end == C.t_end; ++j) if (keys[j] < key) rank += 1;
            dst[lo + rank] = C;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { H->log_count = 0; H->conn_count = 0; }
}

/* ---- host side ---------------------------------------------------------------------------------------------------------- */
/* Engine kernels.  Three designs on one engine source:
 *   a replica per WARP, state staged in shared memory (k_run; k_run_wide for 13-16 resident replicas per SM;
 *   k_run_in_place with the state left in HBM for replicas larger than shared memory);
 *   a replica per LANE, state in place in HBM, each warp running the event kind most of its lanes wait on (k_run_vote,
 *   k_run_vote_hbm);
 *   a replica per LANE, the replicas of a CTA regrouped by the kind of their next event through rings in shared memory
 *   (k_run_pool, k_run_pool_hbm): the default from 8 192 replicas up (4 096 for models read in place), the warp
 *   kernel below that.
 * ITSB_KERNEL=warp | vote | pool overrides (lanes | sorted | phased too in a -DITSB_EXPERIMENTS build: measured slower,
 * profiles/README.md), ITSB_KERNEL_CTAS=n sets the CTAs per SM of k_run_vote. */
enum { K_WARP = 0, K_WIDE, K_IN_PLACE, K_PHASED, K_VOTE, K_LANES, K_SORTED, K_POOL, K_VOTE_HBM, K_POOL_HBM };
static const char* const KERNEL_NAMES[] = {"k_run", "k_run_wide", "k_run_in_place", "k_run_phased", "k_run_vote",
                                           "k_run_lanes", "k_run_sorted", "k_run_pool", "k_run_vote_hbm", "k_run_pool_hbm"};

struct itsb_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaDeviceProp prop;
    itsb_model_desc desc;
    bool loaded = false;
    void* d_tables[ITSB_NUM_BLOBS] = {nullptr};
    RunParams rp;
    StaticOffsets so;
    unsigned char* d_hot = nullptr;
    ArenaNode* d_arena = nullptr;
    uint32_t* d_arena_free = nullptr;
    itsb_trace_rec* d_trace = nullptr;
    itsb_net_event* d_netlog = nullptr;
    itsb_connection* d_connlog = nullptr;
    ConnEnt* d_conntab = nullptr;
    unsigned long long* d_ticket = nullptr;       /* [0] ticket, [1] events_done */
    long long* d_first_failed = nullptr;
    unsigned long long* d_telemetry = nullptr;
    Model* d_model = nullptr;                     /* the table pointers, in device memory (template build) */
    unsigned long long* d_scratch = nullptr;      /* itsb_allreduce_u64 */
    /* streaming drains: per-replica offsets (two arrays of n + 1), two alternating sets of staging buffers */
    struct DrainSet {
        itsb_net_event* d_net = nullptr; itsb_connection* d_conn = nullptr;       /* device, packed */
        itsb_net_event* h_net = nullptr; itsb_connection* h_conn = nullptr;       /* pinned host */
        unsigned long long* h_first = nullptr;                                    /* pinned: net_first | conn_first */
        size_t cap_net = 0, cap_conn = 0, cap_first = 0;
        uint64_t n_net = 0, n_conn = 0, dropped[2] = {0, 0}, bytes = 0;
        cudaEvent_t packed = nullptr, home = nullptr, t0 = nullptr, t1 = nullptr;
        bool pending = false;
    } drain[2];
    unsigned long long* d_first = nullptr;        /* net_first[n + 1] | conn_first[n + 1] | dropped[2] */
    size_t cap_first = 0;
    cudaStream_t copy_stream = nullptr;
    unsigned drain_next = 0;
    uint64_t n_replicas = 0;
    uint64_t launches = 0;
    uint32_t warps_per_cta = 0, grid = 0;
    size_t smem_bytes = 0;
    bool run_pending = false;
    int kernel = 0;                               /* K_*: which engine kernel itsb_run launches (choose_kernel) */
    int kernel_ctas = 3;                          /* CTAs of 256 threads per SM for the lane-per-replica kernels */
    bool forced_in_place = false;
    bool lane_smem_ok = false;                    /* the lane kernels' dynamic shared memory (model tables) was granted */
    int pool_shape = 2;                           /* which instantiation of k_run_pool (itsb_pool.cu POOL_SHAPES) */
    bool pool_ok = false, pool_hbm_ok = false;
    std::string kernel_label;
    float last_ms = 0.f;
    long long first_failed = -1;
    std::string err;
};

static void choose_kernel(itsb_engine* e) {
    const char* k = getenv("ITSB_KERNEL");
    const char* c = getenv("ITSB_KERNEL_CTAS");
    e->kernel_ctas = c && atoi(c) >= 1 && atoi(c) <= 8 ? atoi(c) : 3;
    const std::string want = k ? k : "";
    /* a lane per replica: the model tables staged beside three CTAs per SM (k_run_vote), or left in HBM when they
       exceed shared memory (k_run_vote_hbm: config C4) */
    const bool lane_ok = !e->rp.tables_in_hbm && !e->forced_in_place && e->lane_smem_ok;
    const bool lane_hbm_ok = e->rp.tables_in_hbm && !e->forced_in_place;
    int warp_kernel = e->rp.state_in_hbm ? K_IN_PLACE : (e->warps_per_cta > 12 ? K_WIDE : K_WARP);
#if defined(ITSB_EXPERIMENTS)
    if (want == "phased" && warp_kernel == K_WARP) warp_kernel = K_PHASED;
#endif
    if (lane_ok && want == "vote") e->kernel = K_VOTE;
    else if (lane_hbm_ok && want == "vote") e->kernel = K_VOTE_HBM;
    else if (lane_hbm_ok && e->pool_hbm_ok && want == "pool") e->kernel = K_POOL_HBM;
    else if (lane_ok && e->pool_ok && want == "pool") e->kernel = K_POOL;
#if defined(ITSB_EXPERIMENTS)
    else if (lane_ok && want == "lanes") e->kernel = K_LANES;
    else if (lane_ok && want == "sorted") e->kernel = K_SORTED;
#endif
    else if (want == "warp" || want == "phased" || (!lane_ok && !lane_hbm_ok)) e->kernel = warp_kernel;
    else if (lane_hbm_ok) {
        /* nothing on the shipped-itsim path is lane-parallel, so a lane per replica only needs more replicas than the
           warp kernel holds in flight (148 SMs x 16-20 warps) to win: measured on config C4 (profiles/README.md) */
        /* config C4, 10 000 replicas (x 10^8 events/s): k_run_in_place 0.90, k_run_vote_hbm 1.12, k_run_pool_hbm 1.56 */
        e->kernel = e->n_replicas >= 4096 ? (e->pool_hbm_ok ? K_POOL_HBM : K_VOTE_HBM) : warp_kernel;
    }
    else {
        /* A lane per replica needs replicas to fill the GPU's threads; a warp per replica (k_run, state in shared memory)
           does not.  Measured on the default model, round 2 (x 10^9 events/s; profiles/README.md):
               replicas     k_run   k_run_vote   k_run_pool
                 12 500      1.48      1.51         2.03
                 25 000      1.51      2.12         3.25
                 50 000      1.53      3.10         4.97
                100 000      1.54      4.77         6.09  (7.2 with the later changes to k_run_pool)
                200 000        -         -          6.18
           so the pooled lane kernel from 8 192 replicas up, the warp kernel below.  Models that keep connection records
           (config C5, 20 000 replicas, drained every step): k_run_pool 4.3, k_run 3.4 x 10^8. */
        const bool enough = e->n_replicas >= 8192;
        e->kernel = enough ? (e->pool_ok ? K_POOL : K_VOTE) : warp_kernel;
    }
}


#define CU(e, call)                                                                            \
    do {                                                                                       \
        cudaError_t _rc = (call);                                                              \
        if (_rc != cudaSuccess) {                                                              \
            (e)->err = std::string(#call) + ": " + cudaGetErrorString(_rc);                    \
            return ITSB_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

static void drain_free(itsb_engine* e);
static void free_replicas(itsb_engine* e) {
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    drain_free(e);
    cudaFree(e->d_hot); cudaFree(e->d_arena); cudaFree(e->d_arena_free); cudaFree(e->d_trace); cudaFree(e->d_netlog);
    cudaFree(e->d_connlog); cudaFree(e->d_conntab);
    e->d_connlog = nullptr; e->d_conntab = nullptr;
    e->d_hot = nullptr; e->d_arena = nullptr; e->d_arena_free = nullptr; e->d_trace = nullptr; e->d_netlog = nullptr;
    e->n_replicas = 0;
}

static void drain_free(itsb_engine* e) {
    for (auto& D : e->drain) {
        cudaFree(D.d_net); cudaFree(D.d_conn); cudaFreeHost(D.h_net); cudaFreeHost(D.h_conn); cudaFreeHost(D.h_first);
        if (D.packed) cudaEventDestroy(D.packed);
        if (D.home) cudaEventDestroy(D.home);
        if (D.t0) cudaEventDestroy(D.t0);
        if (D.t1) cudaEventDestroy(D.t1);
        D = itsb_engine::DrainSet();
    }
    cudaFree(e->d_first); e->d_first = nullptr; e->cap_first = 0;
}

template <typename T>
static cudaError_t grow(T*& dev, T*& host, size_t& cap, size_t want) {
    if (want <= cap) return cudaSuccess;
    const size_t n = want + want / 4 + 1024;
    cudaFree(dev); cudaFreeHost(host); dev = nullptr; host = nullptr; cap = 0;
    cudaError_t rc = cudaMalloc(&dev, n * sizeof(T));
    if (rc == cudaSuccess) rc = cudaHostAlloc(&host, n * sizeof(T), cudaHostAllocDefault);
    if (rc == cudaSuccess) cap = n;
    return rc;
}

extern "C" {

int itsb_abi_version(void) { return ITSB_ABI_VERSION; }

int itsb_create(int device, itsb_engine** out) {
    if (!out) return ITSB_ERR_STATE;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return ITSB_ERR_NO_DEVICE;
    }
    itsb_engine* e = new itsb_engine();
    e->device = device;
    *out = e;
    CU(e, cudaSetDevice(device));
    CU(e, cudaGetDeviceProperties(&e->prop, device));
    CU(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CU(e, cudaEventCreate(&e->ev_start));
    CU(e, cudaEventCreate(&e->ev_stop));
    CU(e, cudaMalloc(&e->d_ticket, 2 * sizeof(unsigned long long)));
    CU(e, cudaMalloc(&e->d_first_failed, sizeof(long long)));
    CU(e, cudaMalloc(&e->d_telemetry, ITSB_TM_SIZE * sizeof(unsigned long long)));
    return ITSB_OK;
}

int itsb_load_model(itsb_engine* e, const itsb_model_desc* d, const void* const* blobs, const size_t* sizes) {
    if (!e || !d || !blobs || !sizes) return ITSB_ERR_STATE;
    if (!validate_model(*d, blobs, sizes, e->err)) return ITSB_ERR_BAD_MODEL;
    CU(e, cudaSetDevice(e->device));
    free_replicas(e);
    for (int i = 0; i < ITSB_NUM_BLOBS; ++i) { cudaFree(e->d_tables[i]); e->d_tables[i] = nullptr; }
    e->desc = *d;
    for (int i = 0; i < ITSB_NUM_BLOBS; ++i) {
        size_t padded = (sizes[i] + 15) / 16 * 16 + 16;   /* staged with 128-bit loads */
        CU(e, cudaMalloc(&e->d_tables[i], padded));
        CU(e, cudaMemsetAsync(e->d_tables[i], 0, padded, e->stream));
        if (sizes[i]) CU(e, cudaMemcpyAsync(e->d_tables[i], blobs[i], sizes[i], cudaMemcpyHostToDevice, e->stream));
    }
    CU(e, cudaStreamSynchronize(e->stream));
    memset(&e->rp, 0, sizeof(e->rp));
    Model& m = e->rp.model;
    m.code = (const u64*)e->d_tables[ITSB_BLOB_CODE];
    m.entries = (const itsb_entry*)e->d_tables[ITSB_BLOB_ENTRIES];
    m.dists = (const itsb_dist*)e->d_tables[ITSB_BLOB_DISTS];
    m.consts = (const double*)e->d_tables[ITSB_BLOB_CONSTS];
    m.nets = (const itsb_net*)e->d_tables[ITSB_BLOB_NETS];
    m.nodes = (const itsb_node*)e->d_tables[ITSB_BLOB_NODES];
    m.init = (const itsb_initproc*)e->d_tables[ITSB_BLOB_INIT];
    m.filters = (const itsb_filter*)e->d_tables[ITSB_BLOB_FILTERS];
    m.ifaces = (const itsb_iface*)e->d_tables[ITSB_BLOB_IFACES];
    m.routes = (const itsb_route*)e->d_tables[ITSB_BLOB_ROUTES];
    m.members = (const u32*)e->d_tables[ITSB_BLOB_MEMBERS];
    m.n_ifaces = d->n_ifaces; m.n_routes = d->n_routes; m.n_members = d->n_members;
    m.n_code = d->n_code; m.n_entries = d->n_entries; m.n_dists = d->n_dists; m.n_consts = d->n_consts;
    m.n_nets = d->n_nets; m.n_nodes = d->n_nodes; m.n_init = d->n_init; m.n_filters = d->n_filters;
    if (!e->d_model) CU(e, cudaMalloc(&e->d_model, sizeof(Model)));
    CU(e, cudaMemcpy(e->d_model, &m, sizeof(Model), cudaMemcpyHostToDevice));
    e->rp.layout = compute_layout(*d);
    e->so = static_offsets(*d);
    e->rp.static_bytes = e->so.total;

    /* one CTA per SM; as many resident replicas (warps) as shared memory holds */
    const size_t smem_max = e->prop.sharedMemPerBlockOptin;
    const size_t per_warp = (size_t)e->rp.layout.hot_bytes + 8;
    e->rp.model_dev = e->d_model;
    e->rp.tables_in_hbm = (e->so.total + 4096 > smem_max) ? 1u : 0u;
    if (e->rp.tables_in_hbm) { e->so.total = 128; e->rp.static_bytes = 128; }
    const char* force = getenv("ITSB_STATE_IN_HBM");
    e->rp.state_in_hbm = (e->so.total + per_warp + 128 > smem_max) || (force && force[0] == '1') ? 1u : 0u;
    uint32_t w;
    if (e->rp.state_in_hbm) {
        const char* nw = getenv("ITSB_IN_PLACE_WARPS");
        w = nw ? (uint32_t)atoi(nw) : 16u;        /* no shared-memory limit: registers are (640 threads x 102) */
        if (w < 1 || w > 20) w = 16;
        e->smem_bytes = e->so.total + (size_t)w * 8 + 64;
    } else {
        w = (uint32_t)((smem_max - e->so.total - 128) / per_warp);
        if (w > 16) w = 16;
        const char* cap = getenv("ITSB_MAX_WARPS");      /* experiments: fewer resident replicas per SM */
        if (cap && atoi(cap) >= 1 && (uint32_t)atoi(cap) < w) w = (uint32_t)atoi(cap);
        e->smem_bytes = e->so.total + (size_t)w * e->rp.layout.hot_bytes + (size_t)w * 8 + 64;
    }
    e->warps_per_cta = w;
    e->rp.warps_per_cta = w;
    e->forced_in_place = force && force[0] == '1';
    e->grid = (uint32_t)e->prop.multiProcessorCount;
    CU(e, cudaFuncSetAttribute(k_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
    CU(e, cudaFuncSetAttribute(k_run_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
    CU(e, cudaFuncSetAttribute(k_run_in_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
    /* the lane-per-replica kernel stages only the model tables: more than 48 KB of them (a 4- or 8-subnet LAN: 57 KB,
       115 KB) needs the opt-in too; tables that leave no room for its CTAs send the model to the warp kernels */
    e->lane_smem_ok = !e->rp.tables_in_hbm &&
                      cudaFuncSetAttribute(k_run_vote, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->so.total) == cudaSuccess;
    cudaGetLastError();
    {
        const char* shape = getenv("ITSB_POOL_SHAPE");
        e->pool_shape = shape && atoi(shape) >= 0 && atoi(shape) < itsb_pool_shapes() ? atoi(shape) : 2;
        e->pool_ok = e->lane_smem_ok && itsb_pool_prepare(e->pool_shape, e->so.total);
        e->pool_hbm_ok = e->rp.tables_in_hbm && itsb_pool_prepare_hbm(0, e->so.total);
    }
#if defined(ITSB_EXPERIMENTS)
    CU(e, cudaFuncSetAttribute(k_run_phased, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
    if (e->lane_smem_ok) {
        CU(e, cudaFuncSetAttribute(k_run_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->so.total));
        CU(e, cudaFuncSetAttribute(k_run_sorted, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(e->so.total + 256 * (sizeof(PendingEvent) + sizeof(RunCtx) + 2 + 1 + 4) + 64)));
    }
#endif
    choose_kernel(e);
    e->loaded = true;
    return ITSB_OK;
}

int itsb_init_replicas(itsb_engine* e, uint64_t n, uint64_t seed, uint64_t first_replica_id) {
    if (!e || !e->loaded || n == 0) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    free_replicas(e);
    const Layout& L = e->rp.layout;
    CU(e, cudaMalloc(&e->d_hot, (size_t)n * L.hot_bytes));
    CU(e, cudaMalloc(&e->d_arena, (size_t)n * L.arena_nodes * sizeof(ArenaNode)));
    CU(e, cudaMalloc(&e->d_arena_free, (size_t)n * L.arena_nodes * sizeof(uint32_t)));
    if (L.trace_cap) CU(e, cudaMalloc(&e->d_trace, (size_t)n * L.trace_cap * sizeof(itsb_trace_rec)));
    if (L.log_cap) CU(e, cudaMalloc(&e->d_netlog, (size_t)n * L.log_cap * sizeof(itsb_net_event)));
    e->rp.hot = e->d_hot; e->rp.arena = e->d_arena; e->rp.arena_free = e->d_arena_free; e->rp.trace = e->d_trace;
    if (L.conn_cap) {
        CU(e, cudaMalloc(&e->d_connlog, (size_t)n * conn_log_bytes(L.conn_cap)));
        CU(e, cudaMalloc(&e->d_conntab, (size_t)n * L.max_sock * 2 * L.conn_window * sizeof(ConnEnt)));
    }
    e->rp.netlog = e->d_netlog; e->rp.connlog = e->d_connlog; e->rp.conntab = e->d_conntab;
    e->rp.n_replicas = n; e->rp.seed = seed;
    e->rp.ticket = e->d_ticket; e->rp.events_done = e->d_ticket + 1; e->rp.first_failed = e->d_first_failed;
    /* the t = 0 template: temporaries released on every way out */
    struct Temp {
        void* p[3] = {nullptr, nullptr, nullptr};
        ~Temp() { for (void* q : p) cudaFree(q); }
    } tmp;
    CU(e, cudaMalloc(&tmp.p[0], L.hot_bytes));
    CU(e, cudaMalloc(&tmp.p[1], (size_t)L.arena_nodes * sizeof(ArenaNode)));
    CU(e, cudaMalloc(&tmp.p[2], (size_t)L.arena_nodes * sizeof(uint32_t)));
    unsigned char* tmpl = (unsigned char*)tmp.p[0];
    ArenaNode* tmpl_arena = (ArenaNode*)tmp.p[1];
    uint32_t* tmpl_free = (uint32_t*)tmp.p[2];
    CU(e, cudaMemsetAsync(tmpl, 0, L.hot_bytes, e->stream));
    {
        std::vector<uint32_t> ident(L.arena_nodes);
        for (uint32_t i = 0; i < L.arena_nodes; ++i) ident[i] = L.arena_nodes - 1 - i;
        CU(e, cudaMemcpyAsync(tmpl_free, ident.data(), ident.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
        CU(e, cudaStreamSynchronize(e->stream));
    }
    k_make_template<<<1, 32, 0, e->stream>>>(e->rp, e->d_model, tmpl, tmpl_arena, tmpl_free);
    CU(e, cudaGetLastError());
    uint32_t blocks = (uint32_t)(n < (uint64_t)e->prop.multiProcessorCount * 8 ? n : (uint64_t)e->prop.multiProcessorCount * 8);
    k_fan_out<<<blocks, 256, 0, e->stream>>>(e->rp, tmpl, tmpl_arena, first_replica_id);
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(e->stream));
    const long long none = 0x7FFFFFFFFFFFFFFFll;
    CU(e, cudaMemcpy(e->d_first_failed, &none, sizeof(none), cudaMemcpyHostToDevice));
    CU(e, cudaMemset(e->d_ticket, 0, 2 * sizeof(unsigned long long)));
    e->launches += 2;
    e->n_replicas = n;
    choose_kernel(e);      /* the replica count is part of the choice */
    e->first_failed = -1;
    e->run_pending = false;
    return ITSB_OK;
}

int itsb_run_async(itsb_engine* e, double duration) {
    if (!e || !e->n_replicas) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    /* launches queue on the engine's stream: the ticket is reset per launch, events_done and first_failed
       accumulate until itsb_sync collects them */
    CU(e, cudaMemsetAsync(e->d_ticket, 0, sizeof(unsigned long long), e->stream));
    e->rp.duration = duration;
    if (!e->run_pending) CU(e, cudaEventRecord(e->ev_start, e->stream));
    const uint32_t lane_grid = e->grid * (uint32_t)e->kernel_ctas;
    {   /* lanes per warp that take replicas: all 32 once the batch fills the grid's threads */
        const uint64_t warps = (uint64_t)lane_grid * 8;
        uint64_t lpw = (e->n_replicas + warps - 1) / warps;
        const char* force = getenv("ITSB_VOTE_LANES");
        if (force && atoi(force) >= 1) lpw = (uint64_t)atoi(force);
        e->rp.lanes_per_warp = (uint32_t)(lpw < 1 ? 1 : (lpw > 32 ? 32 : lpw));
        const char* stick = getenv("ITSB_POOL_STICK");
        e->rp.pool_stick = stick ? (uint32_t)atoi(stick) : 4u;     /* measured: 0 -> 6.0, 1..24 -> 6.3-6.7 x 10^9 */
        const char* byprog = getenv("ITSB_POOL_BY_PROGRAM");
        e->rp.pool_by_program = byprog ? (uint32_t)atoi(byprog) : 0u;
    }
    switch (e->kernel) {
    case K_VOTE:     k_run_vote<<<lane_grid, 256, e->so.total, e->stream>>>(e->rp); break;
    case K_VOTE_HBM: k_run_vote_hbm<<<lane_grid, 256, e->so.total, e->stream>>>(e->rp); break;
    case K_POOL:     CU(e, itsb_pool_launch(e->pool_shape, (int)e->grid, e->so.total, e->rp, e->stream)); break;
    case K_POOL_HBM: CU(e, itsb_pool_launch_hbm(0, (int)e->grid, e->so.total, e->rp, e->stream)); break;
#if defined(ITSB_EXPERIMENTS)
    case K_LANES:    k_run_lanes<<<lane_grid, 256, e->so.total, e->stream>>>(e->rp); break;
    case K_SORTED:   k_run_sorted<<<lane_grid, 256, e->so.total + 256 * (sizeof(PendingEvent) + sizeof(RunCtx) + 2 + 1 + 4) + 64,
                                    e->stream>>>(e->rp); break;
    case K_PHASED:   k_run_phased<<<e->grid, e->warps_per_cta * 32, e->smem_bytes, e->stream>>>(e->rp); break;
#endif
    case K_IN_PLACE: k_run_in_place<<<e->grid, e->warps_per_cta * 32, e->smem_bytes, e->stream>>>(e->rp); break;
    case K_WIDE:     k_run_wide<<<e->grid, e->warps_per_cta * 32, e->smem_bytes, e->stream>>>(e->rp); break;
    default:         k_run<<<e->grid, e->warps_per_cta * 32, e->smem_bytes, e->stream>>>(e->rp); break;
    }
    CU(e, cudaGetLastError());
    CU(e, cudaEventRecord(e->ev_stop, e->stream));
    e->launches += 1;
    e->run_pending = true;
    return ITSB_OK;
}

int itsb_sync(itsb_engine* e, uint64_t* events_done) {
    if (!e) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaStreamSynchronize(e->stream));
    if (e->run_pending) {
        e->run_pending = false;
        CU(e, cudaEventElapsedTime(&e->last_ms, e->ev_start, e->ev_stop));
    }
    unsigned long long host[2] = {0, 0};
    long long ff = 0;
    const long long none = 0x7FFFFFFFFFFFFFFFll;
    CU(e, cudaMemcpy(host, e->d_ticket, sizeof(host), cudaMemcpyDeviceToHost));
    CU(e, cudaMemcpy(&ff, e->d_first_failed, sizeof(ff), cudaMemcpyDeviceToHost));
    CU(e, cudaMemset(e->d_ticket + 1, 0, sizeof(unsigned long long)));
    if (events_done) *events_done = host[1];
    if (ff != none) {
        e->first_failed = ff;
        Hdr h;
        CU(e, cudaMemcpy(&h, e->d_hot + (size_t)ff * e->rp.layout.hot_bytes, sizeof(Hdr), cudaMemcpyDeviceToHost));
        return h.status;
    }
    return ITSB_OK;
}

int itsb_run(itsb_engine* e, double duration, uint64_t* events_done) {
    int rc = itsb_run_async(e, duration);
    if (rc) return rc;
    return itsb_sync(e, events_done);
}

int itsb_last_run_ms(itsb_engine* e, float* ms) { if (!e || !ms) return ITSB_ERR_STATE; *ms = e->last_ms; return ITSB_OK; }
const char* itsb_kernel_name(itsb_engine* e) { return e && e->loaded ? KERNEL_NAMES[e->kernel] : ""; }

int itsb_launch_count(itsb_engine* e, uint64_t* n) { if (!e || !n) return ITSB_ERR_STATE; *n = e->launches; return ITSB_OK; }

int itsb_read_now(itsb_engine* e, uint64_t first, uint64_t n, double* dst) {
    if (!e || first + n > e->n_replicas) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    /* now is the first field of each hot block: a strided 2-D copy */
    CU(e, cudaMemcpy2D(dst, sizeof(double), e->d_hot + (size_t)first * e->rp.layout.hot_bytes, e->rp.layout.hot_bytes,
                       sizeof(double), n, cudaMemcpyDeviceToHost));
    return ITSB_OK;
}

int itsb_read_trace(itsb_engine* e, uint64_t replica, itsb_trace_rec* dst, size_t cap, size_t* n_total) {
    if (!e || replica >= e->n_replicas) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    Hdr h;
    CU(e, cudaMemcpy(&h, e->d_hot + (size_t)replica * e->rp.layout.hot_bytes, sizeof(Hdr), cudaMemcpyDeviceToHost));
    if (n_total) *n_total = h.trace_count;
    size_t n = h.trace_count < e->rp.layout.trace_cap ? h.trace_count : e->rp.layout.trace_cap;
    if (n > cap) n = cap;
    if (n && e->d_trace)
        CU(e, cudaMemcpy(dst, e->d_trace + (size_t)replica * e->rp.layout.trace_cap, n * sizeof(itsb_trace_rec), cudaMemcpyDeviceToHost));
    return ITSB_OK;
}

int itsb_read_net_events(itsb_engine* e, uint64_t replica, itsb_net_event* dst, size_t cap, size_t* n_total) {
    if (!e || replica >= e->n_replicas) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    Hdr h;
    CU(e, cudaMemcpy(&h, e->d_hot + (size_t)replica * e->rp.layout.hot_bytes, sizeof(Hdr), cudaMemcpyDeviceToHost));
    if (n_total) *n_total = h.log_count;
    size_t n = h.log_count < e->rp.layout.log_cap ? h.log_count : e->rp.layout.log_cap;
    if (n > cap) n = cap;
    if (n && e->d_netlog)
        CU(e, cudaMemcpy(dst, e->d_netlog + (size_t)replica * e->rp.layout.log_cap, n * sizeof(itsb_net_event), cudaMemcpyDeviceToHost));
    return ITSB_OK;
}

int itsb_read_connections(itsb_engine* e, uint64_t replica, itsb_connection* dst, size_t cap, size_t* n_total) {
    if (!e || replica >= e->n_replicas) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    Hdr h;
    CU(e, cudaMemcpy(&h, e->d_hot + (size_t)replica * e->rp.layout.hot_bytes, sizeof(Hdr), cudaMemcpyDeviceToHost));
    if (n_total) *n_total = h.conn_count;
    size_t n = h.conn_count < e->rp.layout.conn_cap ? h.conn_count : e->rp.layout.conn_cap;
    if (n > cap) n = cap;
    if (n && e->d_connlog) {
        /* records, then their order keys (kept for everything the log holds: a record retired early may belong
           before one that was cut off by `cap`, so order first, cut afterwards) */
        const size_t held = h.conn_count < e->rp.layout.conn_cap ? h.conn_count : e->rp.layout.conn_cap;
        const unsigned char* base = (const unsigned char*)e->d_connlog + (size_t)replica * conn_log_bytes(e->rp.layout.conn_cap);
        std::vector<itsb_connection> recs(held);
        std::vector<uint32_t> keys(held);
        CU(e, cudaMemcpy(recs.data(), base, held * sizeof(itsb_connection), cudaMemcpyDeviceToHost));
        CU(e, cudaMemcpy(keys.data(), base + (size_t)e->rp.layout.conn_cap * sizeof(itsb_connection), held * sizeof(uint32_t),
                         cudaMemcpyDeviceToHost));
        order_connections(recs.data(), keys.data(), held);
        memcpy(dst, recs.data(), n * sizeof(itsb_connection));
    }
    return ITSB_OK;
}

/* histogram of one counter over the replicas: privatised in shared memory, then one atomic per bin and block */
extern "C" __global__ void k_histogram(const unsigned char* hot, uint32_t hot_bytes, uint32_t off_stats, uint64_t n,
                                       uint32_t slot, unsigned long long lo, unsigned long long width, uint32_t nbins,
                                       int only_nonzero, unsigned long long* bins) {
    extern __shared__ unsigned int local[];
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) local[b] = 0;
    __syncthreads();
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = ((const unsigned long long*)(hot + r * hot_bytes + off_stats))[slot];
        if (only_nonzero && v == 0) continue;
        unsigned long long b = v <= lo ? 0 : (v - lo) / width;
        if (b >= nbins) b = nbins - 1;
        atomicAdd(&local[(uint32_t)b], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x)
        if (local[b]) atomicAdd(&bins[b], (unsigned long long)local[b]);
}

static int reduce_telemetry(itsb_engine* e) {
    CU(e, cudaMemsetAsync(e->d_telemetry, 0, ITSB_TM_SIZE * sizeof(unsigned long long), e->stream));
    uint32_t blocks = (uint32_t)(e->n_replicas < 1024 ? e->n_replicas : 1024);
    static_assert(ITSB_TM_SIZE <= 128, "k_reduce_telemetry: one thread per telemetry slot");
    k_reduce_telemetry<<<blocks, 128, 0, e->stream>>>(e->d_hot, e->rp.layout.hot_bytes, e->rp.layout.off_stats, e->n_replicas, e->d_telemetry);
    CU(e, cudaGetLastError());
    e->launches += 1;
    return ITSB_OK;
}

/* ---- itsb_drain_begin / itsb_drain_end ------------------------------------------------------------------------------- */
int itsb_drain_begin(itsb_engine* e) {
    if (!e || !e->n_replicas) return ITSB_ERR_STATE;
    const Layout& L = e->rp.layout;
    if (L.log_cap == 0 && L.conn_cap == 0) { e->err = "the model keeps no record logs"; return ITSB_ERR_STATE; }
    auto& D = e->drain[e->drain_next & 1];
    if (D.pending) { e->err = "two drains are already pending: call itsb_drain_end first"; return ITSB_ERR_STATE; }
    CU(e, cudaSetDevice(e->device));
    const uint64_t n = e->n_replicas;
    if (!e->copy_stream) CU(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    if (!D.packed) {
        CU(e, cudaEventCreateWithFlags(&D.packed, cudaEventDisableTiming));
        CU(e, cudaEventCreateWithFlags(&D.home, cudaEventDisableTiming));
        CU(e, cudaEventCreate(&D.t0));
        CU(e, cudaEventCreate(&D.t1));
    }
    if (e->cap_first < n + 1) {
        cudaFree(e->d_first); e->d_first = nullptr; e->cap_first = 0;
        CU(e, cudaMalloc(&e->d_first, (2 * (n + 1) + 2) * sizeof(unsigned long long)));
        e->cap_first = n + 1;
    }
    if (D.cap_first < n + 1) {
        cudaFreeHost(D.h_first); D.h_first = nullptr; D.cap_first = 0;
        CU(e, cudaHostAlloc(&D.h_first, 2 * (n + 1) * sizeof(unsigned long long), cudaHostAllocDefault));
        D.cap_first = n + 1;
    }
    unsigned long long* net_first = e->d_first;
    unsigned long long* conn_first = e->d_first + (n + 1);
    unsigned long long* dropped = e->d_first + 2 * (n + 1);
    CU(e, cudaEventRecord(D.t0, e->stream));
    CU(e, cudaMemsetAsync(dropped, 0, 2 * sizeof(unsigned long long), e->stream));
    k_log_count<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->d_hot, L.hot_bytes, n, L.log_cap, L.conn_cap,
                                                                     net_first, conn_first, dropped);
    CU(e, cudaGetLastError());
    k_log_scan<<<1, 1024, 0, e->stream>>>(net_first, conn_first, n);
    CU(e, cudaGetLastError());
    e->launches += 2;
    /* the totals size the packed buffers: this is the one wait of a drain (the run before it has to be over anyway) */
    unsigned long long totals[4];
    CU(e, cudaMemcpyAsync(&totals[0], net_first + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(&totals[1], conn_first + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(&totals[2], dropped, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    D.n_net = totals[0]; D.n_conn = totals[1]; D.dropped[0] = totals[2]; D.dropped[1] = totals[3];
    CU(e, grow(D.d_net, D.h_net, D.cap_net, (size_t)D.n_net));
    CU(e, grow(D.d_conn, D.h_conn, D.cap_conn, (size_t)D.n_conn));
    k_log_gather<<<(unsigned)n, 128, 0, e->stream>>>(e->d_hot, L.hot_bytes, n, e->d_netlog, L.log_cap,
                                                      (const unsigned char*)e->d_connlog, L.conn_cap, net_first, conn_first,
                                                      D.d_net, D.d_conn);
    CU(e, cudaGetLastError());
    CU(e, cudaEventRecord(D.t1, e->stream));
    e->launches += 1;
    /* the offsets travel with the records; copied to their own pinned buffer on the engine stream (the next drain
       overwrites the device array), the records on the second stream once the gather is done */
    CU(e, cudaMemcpyAsync(D.h_first, e->d_first, 2 * (n + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaEventRecord(D.packed, e->stream));
    CU(e, cudaStreamWaitEvent(e->copy_stream, D.packed, 0));
    if (D.n_net) CU(e, cudaMemcpyAsync(D.h_net, D.d_net, D.n_net * sizeof(itsb_net_event), cudaMemcpyDeviceToHost, e->copy_stream));
    if (D.n_conn) CU(e, cudaMemcpyAsync(D.h_conn, D.d_conn, D.n_conn * sizeof(itsb_connection), cudaMemcpyDeviceToHost, e->copy_stream));
    CU(e, cudaEventRecord(D.home, e->copy_stream));
    D.bytes = D.n_net * sizeof(itsb_net_event) + D.n_conn * sizeof(itsb_connection) + 2 * (n + 1) * sizeof(unsigned long long) + 32;
    D.pending = true;
    e->drain_next += 1;
    return ITSB_OK;
}

int itsb_drain_end(itsb_engine* e, itsb_drain* out) {
    if (!e || !out) return ITSB_ERR_STATE;
    /* the oldest pending drain */
    itsb_engine::DrainSet* D = nullptr;
    for (unsigned k = 0; k < 2; ++k) {
        auto& C = e->drain[(e->drain_next + k) & 1];
        if (C.pending) { D = &C; break; }
    }
    if (!D) { e->err = "no drain is pending"; return ITSB_ERR_STATE; }
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaEventSynchronize(D->packed));     /* offsets home (engine stream) */
    CU(e, cudaEventSynchronize(D->home));       /* records home (copy stream) */
    const uint64_t n = e->n_replicas;
    out->net = D->h_net; out->n_net = D->n_net; out->net_first = (const uint64_t*)D->h_first;
    out->conn = D->h_conn; out->n_conn = D->n_conn; out->conn_first = (const uint64_t*)(D->h_first + (n + 1));
    out->net_dropped = D->dropped[0]; out->conn_dropped = D->dropped[1];
    out->d2h_bytes = D->bytes;
    float ms = 0.f;
    CU(e, cudaEventElapsedTime(&ms, D->t0, D->t1));
    out->pack_ms = ms;
    D->pending = false;
    return ITSB_OK;
}

int itsb_read_telemetry(itsb_engine* e, uint64_t* dst, size_t n) {
    if (!e || !e->n_replicas || n < ITSB_TM_SIZE) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    int rc = reduce_telemetry(e);
    if (rc) return rc;
    CU(e, cudaMemcpyAsync(dst, e->d_telemetry, ITSB_TM_SIZE * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return ITSB_OK;
}

int itsb_read_replica_telemetry(itsb_engine* e, uint64_t replica, uint64_t* dst, size_t n) {
    if (!e || replica >= e->n_replicas || n < ITSB_TM_SIZE) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaMemcpy(dst, e->d_hot + (size_t)replica * e->rp.layout.hot_bytes + e->rp.layout.off_stats,
                     ITSB_TM_SIZE * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return ITSB_OK;
}

/* ---- NCCL, resolved at run time so the library loads without it ------------------------------------------------------ */
typedef int (*nccl_get_unique_id_t)(void*);
struct nccl_uid { char internal[128]; };
typedef int (*nccl_comm_init_rank_fn)(void**, int, nccl_uid, int);
typedef int (*nccl_comm_destroy_fn)(void*);
typedef int (*nccl_all_reduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static void* g_nccl = nullptr;
static void* nccl_sym(const char* name) {
    if (!g_nccl) g_nccl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl) g_nccl = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return g_nccl ? dlsym(g_nccl, name) : nullptr;
}

int itsb_nccl_unique_id(void* id128, size_t cap) {
    if (cap < 128) return ITSB_ERR_STATE;
    nccl_get_unique_id_t fn = (nccl_get_unique_id_t)nccl_sym("ncclGetUniqueId");
    if (!fn) return ITSB_ERR_NCCL;
    return fn(id128) == 0 ? ITSB_OK : ITSB_ERR_NCCL;
}

int itsb_nccl_comm_init(itsb_engine* e, int nranks, int rank, const void* id128, void** comm_out) {
    if (!e || !id128 || !comm_out) return ITSB_ERR_STATE;
    nccl_comm_init_rank_fn fn = (nccl_comm_init_rank_fn)nccl_sym("ncclCommInitRank");
    if (!fn) { e->err = "libnccl.so.2 not found"; return ITSB_ERR_NCCL; }
    CU(e, cudaSetDevice(e->device));
    nccl_uid uid;
    memcpy(&uid, id128, 128);
    int rc = fn(comm_out, nranks, uid, rank);
    if (rc != 0) { e->err = "ncclCommInitRank failed: " + std::to_string(rc); return ITSB_ERR_NCCL; }
    return ITSB_OK;
}

int itsb_nccl_comm_destroy(void* comm) {
    nccl_comm_destroy_fn fn = (nccl_comm_destroy_fn)nccl_sym("ncclCommDestroy");
    if (!fn) return ITSB_ERR_NCCL;
    return fn(comm) == 0 ? ITSB_OK : ITSB_ERR_NCCL;
}

int itsb_allreduce_telemetry(itsb_engine* e, void* comm, uint64_t* dst, size_t n) {
    if (!e || !e->n_replicas || n < ITSB_TM_SIZE || !comm) return ITSB_ERR_STATE;
    nccl_all_reduce_fn fn = (nccl_all_reduce_fn)nccl_sym("ncclAllReduce");
    if (!fn) { e->err = "libnccl.so.2 not found"; return ITSB_ERR_NCCL; }
    CU(e, cudaSetDevice(e->device));
    int rc = reduce_telemetry(e);
    if (rc) return rc;
    /* ncclUint64 = 5, ncclSum = 0 (nccl.h) */
    int nrc = fn(e->d_telemetry, e->d_telemetry, ITSB_TM_SIZE, 5, 0, comm, e->stream);
    if (nrc != 0) { e->err = "ncclAllReduce failed: " + std::to_string(nrc); return ITSB_ERR_NCCL; }
    CU(e, cudaMemcpyAsync(dst, e->d_telemetry, ITSB_TM_SIZE * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return ITSB_OK;
}

int itsb_allreduce_u64(itsb_engine* e, void* comm, uint64_t* buf, size_t n, int op) {
    if (!e || !comm || !buf || n == 0 || n > 4096 || (op != 0 && op != 2)) return ITSB_ERR_STATE;
    nccl_all_reduce_fn fn = (nccl_all_reduce_fn)nccl_sym("ncclAllReduce");
    if (!fn) { e->err = "libnccl.so.2 not found"; return ITSB_ERR_NCCL; }
    CU(e, cudaSetDevice(e->device));
    if (!e->d_scratch) CU(e, cudaMalloc(&e->d_scratch, 4096 * sizeof(uint64_t)));
    CU(e, cudaMemcpyAsync(e->d_scratch, buf, n * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
    if (fn(e->d_scratch, e->d_scratch, n, 5 /* ncclUint64 */, op, comm, e->stream) != 0) {
        e->err = "ncclAllReduce failed"; return ITSB_ERR_NCCL;
    }
    CU(e, cudaMemcpyAsync(buf, e->d_scratch, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return ITSB_OK;
}

int itsb_histogram(itsb_engine* e, uint32_t slot, uint64_t lo, uint64_t width, uint32_t nbins, int only_nonzero,
                   void* comm, uint64_t* dst) {
    if (!e || !e->n_replicas || !dst || slot >= ITSB_TM_SIZE || width == 0 || nbins == 0 || nbins > 4096) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    unsigned long long* d_bins = nullptr;
    CU(e, cudaMalloc(&d_bins, nbins * sizeof(unsigned long long)));
    cudaError_t c0 = cudaMemsetAsync(d_bins, 0, nbins * sizeof(unsigned long long), e->stream);
    if (c0 == cudaSuccess) {
        const uint32_t blocks = (uint32_t)((e->n_replicas + 255) / 256 < 592 ? (e->n_replicas + 255) / 256 : 592);
        k_histogram<<<blocks, 256, nbins * sizeof(unsigned int), e->stream>>>(
            e->d_hot, e->rp.layout.hot_bytes, e->rp.layout.off_stats, e->n_replicas, slot, lo, width, nbins, only_nonzero, d_bins);
        c0 = cudaGetLastError();
        e->launches += 1;
    }
    if (c0 != cudaSuccess) {
        cudaFree(d_bins);
        e->err = std::string("itsb_histogram: ") + cudaGetErrorString(c0);
        return ITSB_ERR_CUDA;
    }
    int rc = ITSB_OK;
    if (comm) {
        nccl_all_reduce_fn fn = (nccl_all_reduce_fn)nccl_sym("ncclAllReduce");
        if (!fn) { cudaFree(d_bins); e->err = "libnccl.so.2 not found"; return ITSB_ERR_NCCL; }
        if (fn(d_bins, d_bins, nbins, 5 /* ncclUint64 */, 0 /* ncclSum */, comm, e->stream) != 0) {
            e->err = "ncclAllReduce failed"; rc = ITSB_ERR_NCCL;
        }
    }
    if (rc == ITSB_OK) {
        cudaError_t c1 = cudaMemcpyAsync(dst, d_bins, nbins * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream);
        cudaError_t c2 = cudaStreamSynchronize(e->stream);
        if (c1 != cudaSuccess || c2 != cudaSuccess) { e->err = "itsb_histogram: copy failed"; rc = ITSB_ERR_CUDA; }
    }
    cudaFree(d_bins);
    return rc;
}


int itsb_replica_status(itsb_engine* e, uint64_t replica, int32_t* status, uint64_t* detail) {
    if (!e || replica >= e->n_replicas) return ITSB_ERR_STATE;
    CU(e, cudaSetDevice(e->device));
    Hdr h;
    CU(e, cudaMemcpy(&h, e->d_hot + (size_t)replica * e->rp.layout.hot_bytes, sizeof(Hdr), cudaMemcpyDeviceToHost));
    if (status) *status = h.status;
    if (detail) *detail = h.detail;
    return ITSB_OK;
}

int64_t itsb_first_failed_replica(itsb_engine* e) { return e ? e->first_failed : -1; }

int itsb_state_bytes(itsb_engine* e, uint64_t* hot, uint64_t* arena) {
    if (!e || !e->loaded) return ITSB_ERR_STATE;
    if (hot) *hot = e->rp.layout.hot_bytes;
    if (arena) *arena = (uint64_t)e->rp.layout.arena_nodes * (sizeof(ArenaNode) + 4);
    return ITSB_OK;
}

int itsb_eval_log(itsb_engine* e, const double* x, double* y, size_t n) {
    if (!e || !x || !y) return ITSB_ERR_STATE;
    if (n == 0) return ITSB_OK;
    CU(e, cudaSetDevice(e->device));
    double* d = nullptr;
    CU(e, cudaMalloc(&d, n * sizeof(double)));
    cudaError_t rc = cudaMemcpyAsync(d, x, n * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (rc == cudaSuccess) {
        k_eval_log<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(d, n);
        rc = cudaGetLastError();
        e->launches += 1;
    }
    if (rc == cudaSuccess) rc = cudaMemcpyAsync(y, d, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (rc == cudaSuccess) rc = cudaStreamSynchronize(e->stream);
    cudaFree(d);
    if (rc != cudaSuccess) { e->err = std::string("itsb_eval_log: ") + cudaGetErrorString(rc); return ITSB_ERR_CUDA; }
    return ITSB_OK;
}

const char* itsb_last_error(itsb_engine* e) { return e ? e->err.c_str() : ""; }

void itsb_destroy(itsb_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    free_replicas(e);         /* waits for the copy stream, releases the drain buffers */
    if (e->copy_stream) { cudaStreamDestroy(e->copy_stream); e->copy_stream = nullptr; }
    for (int i = 0; i < ITSB_NUM_BLOBS; ++i) cudaFree(e->d_tables[i]);
    cudaFree(e->d_ticket); cudaFree(e->d_first_failed); cudaFree(e->d_telemetry); cudaFree(e->d_model); cudaFree(e->d_scratch);
    if (e->ev_start) cudaEventDestroy(e->ev_start);
    if (e->ev_stop) cudaEventDestroy(e->ev_stop);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

}  // extern "C"
