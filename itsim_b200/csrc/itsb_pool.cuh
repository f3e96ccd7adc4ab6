/*
 * itsb_pool.cuh -- k_run_pool (instantiated by itsb_pool.cu, and by itsb_pool_hbm.cu for models whose tables stay in HBM): a replica per lane, replicas changing warps through per-kind queues in shared memory.
 *
 * k_run_vote regroups inside one warp: the warp runs the event kind most of its own 32 lanes wait on (6 lanes active on
 * average).  Here a CTA's replicas are not tied to lanes at all.  Each replica sits in a *slot* in shared memory (its
 * address, the run's bookkeeping, its pending event: 64 bytes); a slot whose replica is ready sits in the ring of its
 * pending event's kind.  A warp takes up to 32 slots from the fullest ring, its lanes handle those events -- one kind, so
 * one handler -- take each replica's next event and push the slot onto the ring of that event's kind.  No CTA barrier:
 * the rings are the only synchronisation (atomic reservation of positions, entries published with the lap of their
 * position and cleared by the lane that takes them).  A replica still sees its own events in its own order, so results are
 * bit-identical to the other kernels'.
 *
 * Round 1 built this as an experiment and measured it slower than the per-warp vote (2.53 vs 3.75 x 10^9 events/s): the
 * engine then waited on chains of dependent loads, and grouping does not shorten those.  Round 2 shortened them
 * (delivery digests, the event heap, inline queue heads: DESIGN.md section 4), which left k_run_vote bound by the rate
 * at which its SMs can fetch instructions (GPC instruction-cache requests at 91 % of their peak) -- and what grouping
 * buys is exactly that: the same events in fewer warp instructions.  Shapes (threads x CTAs per SM, slots per CTA) are
 * template parameters; itsb_pool_launch picks one (ITSB_POOL_SHAPE overrides; measurements in profiles/README.md).
 */
#pragma once
#define ITSB_LANES 1
#include "itsb_kernel.cuh"

#ifndef POOL_ENTRY            /* the including translation unit names its copies of the host entry points */
#define POOL_ENTRY(name) name
#endif

/* ring index: the event kind (0..8); activations of a parked process (WAKE: the most frequent and the most divergent
 * kind -- a daemon's receive loop, a query's reply, the DHCP exchange run different code) are split further by the
 * program they resume, so that the lanes of a batch run the same code and not just the same handler: measured active
 * lanes per warp instruction 7.4 -> see profiles/README.md */
#ifndef POOL_PROG_RINGS
#define POOL_PROG_RINGS 0       /* measured with 16 (default bench): 6.31 vs 6.6-6.7 x 10^9 with the nine kind rings alone when
                                   one lane walked all the rings, and 6.6-6.8 vs 7.2-7.3 with the warp-wide pick: 25 rings
                                   mean smaller batches and more changes of code per SM, which cost more than the finer
                                   grouping saves; kept as a build option (-DPOOL_PROG_RINGS=16, ITSB_POOL_BY_PROGRAM=1) */
#endif
#define POOL_RINGS (9 + POOL_PROG_RINGS)

__device__ __forceinline__ unsigned int pool_key(Rep& R, const PendingEvent& E, unsigned int by_program) {
    const unsigned int kind = ev_kind(E.data);
#if POOL_PROG_RINGS > 0
    if (by_program != 0 && kind == ITSB_EV_WAKE) return 9u + (R.pcb[ev_idx(E.data)].prog & (POOL_PROG_RINGS - 1));
#endif
    (void)R; (void)by_program;
    return kind;
}

struct PoolSlot { PendingEvent E; RunCtx C; unsigned long long r; };      /* 64 bytes */

template <int SLOTS, int RING>
struct PoolShared {
    PoolSlot slot[SLOTS];
    unsigned int ring[POOL_RINGS][RING];            /* (lap << 16) | (slot + 1); 0 = free */
    unsigned int head[POOL_RINGS];
    unsigned int tail[POOL_RINGS];
    unsigned int live;                              /* slots that hold an unfinished replica */
    unsigned int cur_kind;                          /* the kind the CTA's warps took last (p.pool_stick) */
};

/* A ring cell holds (lap << 16) | (slot + 1), lap = 1 + (position / ring size) mod 32768, or 0 when free.  The producer
 * of a position waits for the cell to be free (the entry of the previous lap taken), the consumer of a position waits
 * for ITS lap's entry: an entry of the previous lap that its own consumer has claimed but not read yet can never be
 * mistaken for the new one. */
template <int RING>
__device__ __forceinline__ unsigned int pool_lap(unsigned int pos) { return ((pos / RING) & 0x7FFFu) + 1u; }

template <int RING>
__device__ __forceinline__ void pool_push(unsigned int* ring, unsigned int* tail, unsigned int slot) {
    const unsigned int pos = atomicAdd(tail, 1u);
    unsigned int* cell = ring + (pos & (RING - 1));
    const unsigned int val = (pool_lap<RING>(pos) << 16) | (slot + 1u);
    while (atomicCAS(cell, 0u, val) != 0u) { }
}

/* binds the views of replica r and starts its run(); returns false if the run is over before its first event */
__device__ __forceinline__ bool pool_begin(const RunParams& p, const Model* m, unsigned long long r, PoolSlot& s,
                                           unsigned int& kind) {
    Rep R; Binding B;
    B.arena = p.arena + r * p.layout.arena_nodes;
    B.arena_free = p.arena_free + r * p.layout.arena_nodes;
    B.trace = p.trace ? p.trace + r * p.layout.trace_cap : nullptr;
    B.netlog = p.netlog ? p.netlog + r * p.layout.log_cap : nullptr;
    B.connlog = p.connlog ? (itsb_connection*)((unsigned char*)p.connlog + r * conn_log_bytes(p.layout.conn_cap)) : nullptr;
    B.conntab = p.conntab ? p.conntab + r * ((size_t)p.layout.max_sock * 2 * p.layout.conn_window) : nullptr;
    bind_views(R, p.hot + r * p.layout.hot_bytes, &p.layout, m, B, p.seed);
    RunCtx C;
    run_begin(R, p.duration, C);
    PendingEvent E;
    s.r = r;
    s.C = C;
    if (run_take(R, C, E)) { s.E = E; kind = pool_key(R, E, p.pool_by_program); return true; }
    atomicAdd(p.events_done, run_events(R, C));
    if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
    return false;
}

template <int THREADS, int MIN_CTAS, int SLOTS, int RING>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_run_pool(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    typedef PoolShared<SLOTS, RING> Shared;
    const Model* m = p.tables_in_hbm ? p.model_dev : stage_model(p.model, smem);
    Shared* S = (Shared*)(smem + (p.tables_in_hbm ? 0 : p.static_bytes));
    const unsigned int t = threadIdx.x, lane = t & 31;
    for (unsigned int i = t; i < POOL_RINGS * RING; i += THREADS) (&S->ring[0][0])[i] = 0;
    if (t < POOL_RINGS) { S->head[t] = 0; S->tail[t] = 0; }
    if (t == 0) { S->live = 0; S->cur_kind = 0; }
    __syncthreads();
    /* the batch is spread evenly over the CTAs: each seeds its share of slots (p.lanes_per_warp carries the share), and a
       slot whose replica finishes takes the next one of the batch */
    const unsigned int my_slots = p.lanes_per_warp < (unsigned int)SLOTS ? p.lanes_per_warp : (unsigned int)SLOTS;
    for (unsigned int sl = t; sl < my_slots; sl += THREADS) {
        unsigned int kind = 0;
        bool ready = false;
        unsigned long long r = atomicAdd(p.ticket, 1ull);
        while (r < p.n_replicas && !(ready = pool_begin(p, m, r, S->slot[sl], kind))) r = atomicAdd(p.ticket, 1ull);
        if (ready) { atomicAdd(&S->live, 1u); __threadfence_block(); pool_push<RING>(S->ring[kind], &S->tail[kind], sl); }
    }
    __syncthreads();
    /* ---- warps pull slots of one kind, handle, push them on ------------------------------------------------------------ */
    unsigned int idle_ns = 32;          /* back-off of a warp that finds every ring empty (doubles up to 2 us) */
    for (;;) {
        unsigned int kind = 0, first = 0, n = 0;
        /* the whole warp picks a ring: lane k reads ring k's fill, one max-reduction finds the fullest (the lowest kind on a
           tie); lane 0 alone claims the slots.  (A single lane walking the rings was 7 % of all executed instructions and
           ~500 cycles of dependent shared-memory loads per batch.) */
        static_assert(POOL_RINGS <= 32, "one lane per ring");
        for (;;) {
            unsigned int avail = 0;
            if (lane < POOL_RINGS) {
                avail = *(volatile unsigned int*)&S->tail[lane] - *(volatile unsigned int*)&S->head[lane];
                if (avail > (unsigned int)RING) avail = 0;
            }
            unsigned int best = 0, best_n = 0;
            if (p.pool_stick != 0) {                /* stay with the kind the other warps are running, while it has work */
                const unsigned int ck = *(volatile unsigned int*)&S->cur_kind;
                const unsigned int a = __shfl_sync(0xFFFFFFFFu, avail, ck);
                if (a >= p.pool_stick) { best = ck; best_n = a; }
            }
            if (best_n == 0) {
                const unsigned int top = __reduce_max_sync(0xFFFFFFFFu, (avail << 5) | (31u - lane));
                best_n = top >> 5;
                best = 31u - (top & 31u);
                if (best_n != 0 && p.pool_stick != 0 && lane == 0) *(volatile unsigned int*)&S->cur_kind = best;
            }
            if (best_n == 0) {
                if (*(volatile unsigned int*)&S->live == 0) { n = 0xFFFFFFFFu; break; }      /* the CTA's work is done */
                /* nothing is ready: every replica of the CTA is in some other warp's hands (a batch smaller than the
                   CTA's lanes: config C4, strong scaling).  Polling the rings flat out costs the working warps issue slots
                   and instruction fetches (large-LAN capture: 44 % of all executed instructions were this loop), so the
                   wait backs off */
                __nanosleep(idle_ns);
                if (idle_ns < 2048) idle_ns <<= 1;
                continue;
            }
            idle_ns = 32;
            if (lane == 0) {
                const unsigned int h = *(volatile unsigned int*)&S->head[best];
                const unsigned int a = *(volatile unsigned int*)&S->tail[best] - h;
                if (a != 0 && a <= (unsigned int)RING) {
                    const unsigned int take = a < 32u ? a : 32u;
                    if (atomicCAS(&S->head[best], h, h + take) == h) { first = h; n = take; }
                }
            }
            n = __shfl_sync(0xFFFFFFFFu, n, 0);
            if (n != 0) { kind = best; first = __shfl_sync(0xFFFFFFFFu, first, 0); break; }
        }
        if (n == 0xFFFFFFFFu) break;
        if (lane < n) {
            volatile unsigned int* cell = &S->ring[kind][(first + lane) & (RING - 1)];
            const unsigned int want = pool_lap<RING>(first + lane);
            unsigned int v;
            while (((v = *cell) >> 16) != want) { }
            *cell = 0u;
            __threadfence_block();              /* the slot's contents were written before the cell was published */
            const unsigned int s = (v & 0xFFFFu) - 1u;
            PoolSlot& SL = S->slot[s];
            unsigned long long r = SL.r;
            Rep R;
            rebind(R, (Hdr*)(p.hot + r * p.layout.hot_bytes));
            run_handle(R, SL.E);
            PendingEvent E;
            const RunCtx C = SL.C;
            if (run_take(R, C, E)) {
                SL.E = E;
                __threadfence_block();
                const unsigned int k2 = pool_key(R, E, p.pool_by_program);
                pool_push<RING>(S->ring[k2], &S->tail[k2], s);
            } else {
                atomicAdd(p.events_done, run_events(R, C));
                if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
                /* the slot takes the next replica of the batch, if there is one */
                unsigned int k2 = 0;
                bool ready = false;
                r = atomicAdd(p.ticket, 1ull);
                while (r < p.n_replicas && !(ready = pool_begin(p, m, r, SL, k2))) r = atomicAdd(p.ticket, 1ull);
                if (ready) { __threadfence_block(); pool_push<RING>(S->ring[k2], &S->tail[k2], s); }
                else atomicSub(&S->live, 1u);
            }
        }
        __syncwarp();
    }
}

/* ---- shapes ------------------------------------------------------------------------------------------------------------- */
struct PoolShape { int threads, ctas_per_sm, slots; size_t shared; void (*kernel)(const RunParams); const char* name; };

#define POOL_SHAPE(T, C, S, R) {T, C, S, sizeof(PoolShared<S, R>) + 64, k_run_pool<T, C, S, R>, #T "x" #C "/" #S}
/* default bench, 100 000 replicas, x 10^9 events/s (profiles/README.md): 256x3/256 5.31, 768x1/768 5.57, 512x1/768 6.09,
   384x2/384 5.32, 512x1/1024 6.10, 384x1/768 5.74, 640x1/768 5.95, 576x1/768 5.83, 448x1/768 5.95 */
static const PoolShape POOL_SHAPES[] = {
#if defined(POOL_ALL_SHAPES)
    POOL_SHAPE(256, 3, 256, 256),       /* round 1's shape */
    POOL_SHAPE(768, 1, 768, 1024),
    POOL_SHAPE(512, 1, 768, 1024),      /* the default (ITSB_POOL_SHAPE=2) */
    POOL_SHAPE(384, 2, 384, 512),
    POOL_SHAPE(512, 1, 1024, 1024),
    POOL_SHAPE(384, 1, 768, 1024),
#else
    POOL_SHAPE(512, 1, 768, 1024),
#endif
};
static const int N_POOL_SHAPES = (int)(sizeof(POOL_SHAPES) / sizeof(POOL_SHAPES[0]));

extern "C" __attribute__((visibility("hidden"))) int POOL_ENTRY(itsb_pool_shapes)(void) { return N_POOL_SHAPES; }
extern "C" __attribute__((visibility("hidden"))) const char* POOL_ENTRY(itsb_pool_shape_name)(int i) { return POOL_SHAPES[i].name; }

/* opt-in shared memory for one shape (tables + slots + rings); returns false when it does not fit */
extern "C" __attribute__((visibility("hidden"))) bool POOL_ENTRY(itsb_pool_prepare)(int shape, size_t table_bytes) {
    const PoolShape& P = POOL_SHAPES[shape];
    const bool ok = cudaFuncSetAttribute(P.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(table_bytes + P.shared)) == cudaSuccess;
    cudaGetLastError();
    return ok;
}

/* Launches shape `shape` over n_sms SMs.  rp.lanes_per_warp is set to the slots a CTA seeds: the batch spread evenly
 * over the CTAs, capped by the shape's slots. */
extern "C" __attribute__((visibility("hidden"))) cudaError_t POOL_ENTRY(itsb_pool_launch)(int shape, int n_sms, size_t table_bytes, RunParams rp,
                                                                            cudaStream_t stream) {
    const PoolShape& P = POOL_SHAPES[shape];
    const unsigned int grid = (unsigned int)(n_sms * P.ctas_per_sm);
    unsigned long long share = (rp.n_replicas + grid - 1) / grid;
    rp.lanes_per_warp = (uint32_t)(share < 1 ? 1 : (share > (unsigned long long)P.slots ? (unsigned long long)P.slots : share));
    P.kernel<<<grid, P.threads, table_bytes + P.shared, stream>>>(rp);
    return cudaGetLastError();
}
