/*
 * itsb_sorted.cu -- EXPERIMENT (ITSB_SORTED=n): a replica per lane, lanes regrouped every step by what their replica
 * does next.
 *
 * k_run_lanes showed that the engine at width 1 keeps up with the warp-per-replica kernel even though only 3 of a
 * warp's 32 lanes are active on average (32 replicas at 32 different places of the code).  Here a CTA owns 256
 * replicas at a time (state in place in HBM) and works in steps: every replica has one pending event, taken from its
 * queue already; the CTA sorts its 256 replicas by the key of that event (kind, program of the process it activates:
 * a counting sort in shared memory), thread i then handles the event of the i-th replica in that order and takes the
 * next one.  A warp therefore runs one handler for 32 replicas instead of 32 handlers for one each.  Every replica sees
 * its own events in its own order, so results are bit-identical to k_run's.
 */
#define ITSB_LANES 1
#include "itsb_kernel.cuh"

#define SORT_CTA 256
#ifndef ITSB_SORTED_MIN_CTAS
#define ITSB_SORTED_MIN_CTAS 3
#endif

extern "C" __global__ void __launch_bounds__(SORT_CTA, ITSB_SORTED_MIN_CTAS) k_run_sorted(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Model* m = p.tables_in_hbm ? p.model_dev : stage_model(p.model, smem);
    unsigned char* base = smem + (p.tables_in_hbm ? 0 : p.static_bytes);
    PendingEvent* pend = (PendingEvent*)base;                                 /* [256] pending event of slot s      */
    RunCtx* ctx = (RunCtx*)(pend + SORT_CTA);                                 /* [256] run() bookkeeping of slot s  */
    unsigned short* order = (unsigned short*)(ctx + SORT_CTA);                /* [256] slots sorted by key          */
    unsigned char* key = (unsigned char*)(order + SORT_CTA);                  /* [256] key of slot s (255: no work) */
    unsigned int* hist = (unsigned int*)(key + SORT_CTA);                     /* [256] counting-sort bins           */
    __shared__ unsigned long long first;
    __shared__ unsigned int live_slots;
    const unsigned int t = threadIdx.x;
    for (;;) {
        if (t == 0) first = atomicAdd(p.ticket, (unsigned long long)SORT_CTA);
        __syncthreads();
        const unsigned long long r0 = first;
        if (r0 >= p.n_replicas) break;
        /* ---- run_begin + first take, slot t = replica r0 + t ------------------------------------------------------- */
        {
            const unsigned long long r = r0 + t;
            key[t] = 255;
            if (r < p.n_replicas) {
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
                if (run_take(R, C, E)) { key[t] = (unsigned char)pending_key(R, E); pend[t] = E; }
                else { atomicAdd(p.events_done, run_events(R, C)); if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r); }
                ctx[t] = C;
            }
        }
        /* ---- steps: sort the slots by key, handle one event per live slot, take the next ------------------------------- */
        for (;;) {
            hist[t] = 0;
            __syncthreads();
            const unsigned int k = key[t];
            atomicAdd(&hist[k], 1u);
            __syncthreads();
            if (t == 0) {                       /* exclusive prefix over 256 bins (serial: 256 adds per step per CTA) */
                unsigned int acc = 0;
                for (int b = 0; b < 256; ++b) { const unsigned int c = hist[b]; hist[b] = acc; acc += c; }
                live_slots = hist[255];                      /* slots with work come first (keys < 255) */
            }
            __syncthreads();
            const unsigned int live = live_slots;
            order[atomicAdd(&hist[k], 1u)] = (unsigned short)t;
            __syncthreads();
            if (live == 0) break;
            if (t < live) {
                const unsigned int s = order[t];
                const unsigned long long r = r0 + s;
                Rep R;
                rebind(R, (Hdr*)(p.hot + r * p.layout.hot_bytes));
                run_handle(R, pend[s]);
                PendingEvent E;
                const RunCtx C = ctx[s];
                if (run_take(R, C, E)) { key[s] = (unsigned char)pending_key(R, E); pend[s] = E; }
                else {
                    key[s] = 255;
                    atomicAdd(p.events_done, run_events(R, C));
                    if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}
