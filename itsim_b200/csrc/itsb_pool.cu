/*
 * itsb_pool.cu -- EXPERIMENT (ITSB_KERNEL=pool): a replica per lane, replicas changing warps through per-kind queues in
 * shared memory.
 *
 * k_run_vote regroups inside one warp: the warp runs the event kind most of its own 32 lanes wait on (6 lanes active on
 * average).  Here a CTA's 256 replicas are not tied to lanes at all.  Each replica sits in a *slot* in shared memory (its
 * address, the run's bookkeeping, its pending event: 64 bytes); a slot whose replica is ready sits in the ring of its
 * pending event's kind.  A warp takes up to 32 slots from the fullest ring, its lanes handle those events -- one kind, so
 * one handler -- take each replica's next event and push the slot onto the ring of that event's kind.  No CTA barrier:
 * the rings are the only synchronisation (atomic reservation of positions, entries published with the lap of their
 * position and cleared by the lane that takes them).  A replica still sees its own events in its own order, so results are
 * bit-identical to the other kernels'.
 */
#define ITSB_LANES 1
#include "itsb_kernel.cuh"

#ifndef POOL_SLOTS
#define POOL_SLOTS 256                  /* replicas a CTA holds at a time (measured: 256 with 3 CTAs per SM 2.53, 1024 with 2 CTAs
                                           1.61, 1024 with 1 CTA 1.37 x 10^9 events/s on the default bench) */
#endif
#define POOL_THREADS 256
#define POOL_RINGS 9                    /* ring index = event kind (0..8) */
#ifndef ITSB_POOL_MIN_CTAS
#define ITSB_POOL_MIN_CTAS 3
#endif

struct PoolSlot { PendingEvent E; RunCtx C; unsigned long long r; };      /* 64 bytes */

struct PoolShared {
    PoolSlot slot[POOL_SLOTS];
    unsigned int ring[POOL_RINGS][POOL_SLOTS];      /* slot + 1; 0 = position reserved but not published yet / free */
    unsigned int head[POOL_RINGS];
    unsigned int tail[POOL_RINGS];
    unsigned int live;                              /* slots that hold an unfinished replica */
};

/* A ring cell holds (lap << 16) | (slot + 1), lap = 1 + (position / ring size) mod 65535, or 0 when free.  The producer
 * of a position waits for the cell to be free (the entry of the previous lap taken), the consumer of a position waits
 * for ITS lap's entry: an entry of the previous lap that its own consumer has claimed but not read yet can never be
 * mistaken for the new one. */
__device__ __forceinline__ unsigned int pool_lap(unsigned int pos) { return (pos / POOL_SLOTS) % 0xFFFFu + 1u; }

__device__ __forceinline__ void pool_push(PoolShared* S, unsigned int kind, unsigned int slot) {
    const unsigned int pos = atomicAdd(&S->tail[kind], 1u);
    volatile unsigned int* cell = &S->ring[kind][pos & (POOL_SLOTS - 1)];
    const unsigned int val = (pool_lap(pos) << 16) | (slot + 1u);
    while (atomicCAS((unsigned int*)cell, 0u, val) != 0u) { }
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
    if (run_take(R, C, E)) { s.E = E; kind = ev_kind(E.data); return true; }
    atomicAdd(p.events_done, run_events(R, C));
    if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
    return false;
}

extern "C" __global__ void __launch_bounds__(POOL_THREADS, ITSB_POOL_MIN_CTAS) k_run_pool(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Model* m = p.tables_in_hbm ? p.model_dev : stage_model(p.model, smem);
    PoolShared* S = (PoolShared*)(smem + (p.tables_in_hbm ? 0 : p.static_bytes));
    const unsigned int t = threadIdx.x, lane = t & 31;
    for (unsigned int i = t; i < POOL_RINGS * POOL_SLOTS; i += POOL_THREADS) (&S->ring[0][0])[i] = 0;
    if (t < POOL_RINGS) { S->head[t] = 0; S->tail[t] = 0; }
    if (t == 0) S->live = 0;
    __syncthreads();
    /* every thread seeds its share of the slots with replicas */
    for (unsigned int sl = t; sl < POOL_SLOTS; sl += POOL_THREADS) {
        unsigned int kind = 0;
        bool ready = false;
        unsigned long long r = atomicAdd(p.ticket, 1ull);
        while (r < p.n_replicas && !(ready = pool_begin(p, m, r, S->slot[sl], kind))) r = atomicAdd(p.ticket, 1ull);
        if (ready) { atomicAdd(&S->live, 1u); pool_push(S, kind, sl); }
    }
    __syncthreads();
    /* ---- warps pull slots of one kind, handle, push them on ------------------------------------------------------------ */
    for (;;) {
        unsigned int kind = 0, first = 0, n = 0;
        if (lane == 0) {
            for (;;) {
                unsigned int best = 0, best_n = 0;
#pragma unroll
                for (unsigned int k = 0; k < POOL_RINGS; ++k) {
                    const unsigned int avail = *(volatile unsigned int*)&S->tail[k] - *(volatile unsigned int*)&S->head[k];
                    if (avail > best_n && avail <= POOL_SLOTS) { best_n = avail; best = k; }
                }
                if (best_n == 0) {
                    if (*(volatile unsigned int*)&S->live == 0) { n = 0xFFFFFFFFu; break; }      /* the CTA's work is done */
                    __nanosleep(200);
                    continue;
                }
                const unsigned int h = *(volatile unsigned int*)&S->head[best];
                const unsigned int avail = *(volatile unsigned int*)&S->tail[best] - h;
                if (avail == 0 || avail > POOL_SLOTS) continue;
                const unsigned int take = avail < 32u ? avail : 32u;
                if (atomicCAS(&S->head[best], h, h + take) == h) { kind = best; first = h; n = take; break; }
            }
        }
        kind = __shfl_sync(0xFFFFFFFFu, kind, 0);
        first = __shfl_sync(0xFFFFFFFFu, first, 0);
        n = __shfl_sync(0xFFFFFFFFu, n, 0);
        if (n == 0xFFFFFFFFu) break;
        if (lane < n) {
            volatile unsigned int* cell = &S->ring[kind][(first + lane) & (POOL_SLOTS - 1)];
            const unsigned int want = pool_lap(first + lane);
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
                pool_push(S, ev_kind(E.data), s);
            } else {
                atomicAdd(p.events_done, run_events(R, C));
                if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
                /* the slot takes the next replica of the batch, if there is one */
                unsigned int k2 = 0;
                bool ready = false;
                r = atomicAdd(p.ticket, 1ull);
                while (r < p.n_replicas && !(ready = pool_begin(p, m, r, SL, k2))) r = atomicAdd(p.ticket, 1ull);
                if (ready) { __threadfence_block(); pool_push(S, k2, s); }
                else atomicSub(&S->live, 1u);
            }
        }
        __syncwarp();
    }
}

extern "C" __attribute__((visibility("hidden"))) size_t itsb_pool_shared_bytes(void) { return sizeof(PoolShared) + 64; }
