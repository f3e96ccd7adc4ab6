/*
 * itsb_vote.cuh -- the lane-per-replica engine kernel's body, shared by its two instantiations:
 *   k_run_vote      (itsb_lanes.cu)      model tables staged in shared memory: the flagship path
 *   k_run_vote_hbm  (itsb_lanes_hbm.cu)  model tables read in place (L1/L2-cached) for models whose tables exceed shared
 *                                        memory (config C4: the 1 040-node LAN), with select()'s helper hops taken
 *                                        without the interpreter (ITSB_HIDDEN_FAST), like k_run_in_place
 * Two translation units so that each kernel's code is laid out on its own.
 */
#pragma once
#define ITSB_LANES 1
#include "itsb_kernel.cuh"

#ifndef ITSB_LANES_MIN_CTAS
#define ITSB_LANES_MIN_CTAS 3
#endif

/* k_run_vote (ITSB_LANES=-n): the same replica per lane, with the warp deciding what to run.  Every lane holds its
 * replica's next event, already taken from the queue; each round the warp picks the key (event kind, program) that most
 * of its lanes are waiting on and only those lanes handle their event and take the next one.  The others keep theirs
 * and become the majority later, so a handler runs for several replicas at a time instead of once per replica. */
__device__ __forceinline__ void run_vote(const RunParams& p, unsigned char* smem) {
    const Model* m = p.tables_in_hbm ? p.model_dev : stage_model(p.model, smem);
    Rep R;
    RunCtx C;
    PendingEvent E;
    unsigned long long r = 0;
    unsigned int key = 0xFFFFu;         /* 0xFFFF: this lane has no replica */
    /* A batch that does not fill the GPU's threads is spread over all warps: with r replicas per warp instead of 32 in
       a few warps, a replica waits for fewer others whose next event is of another kind, and every SM has loads in
       flight (strong scaling: 12 500 replicas per GPU when config C2 is split eight ways). */
    bool exhausted = (threadIdx.x & 31) >= p.lanes_per_warp;
    for (;;) {
        if (key == 0xFFFFu && !exhausted) {
            r = atomicAdd(p.ticket, 1ull);
            if (r >= p.n_replicas) {
                exhausted = true;
            } else {
                Binding B;
                B.arena = p.arena + r * p.layout.arena_nodes;
                B.arena_free = p.arena_free + r * p.layout.arena_nodes;
                B.trace = p.trace ? p.trace + r * p.layout.trace_cap : nullptr;
                B.netlog = p.netlog ? p.netlog + r * p.layout.log_cap : nullptr;
                B.connlog = p.connlog ? (itsb_connection*)((unsigned char*)p.connlog + r * conn_log_bytes(p.layout.conn_cap)) : nullptr;
                B.conntab = p.conntab ? p.conntab + r * ((size_t)p.layout.max_sock * 2 * p.layout.conn_window) : nullptr;
                bind_views(R, p.hot + r * p.layout.hot_bytes, &p.layout, m, B, p.seed);
                run_begin(R, p.duration, C);
                if (run_take(R, C, E)) key = pending_key(R, E);
                else {
                    atomicAdd(p.events_done, run_events(R, C));
                    if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
                }
            }
        }
        __syncwarp();
        const unsigned int have = __ballot_sync(0xFFFFFFFFu, key != 0xFFFFu);
        if (have == 0) {
            if (__all_sync(0xFFFFFFFFu, exhausted)) break;
            continue;
        }
        /* the key most lanes wait on (ties: the larger key) */
        const unsigned int peers = __match_any_sync(0xFFFFFFFFu, key);
        const unsigned int votes = key == 0xFFFFu ? 0u : (((unsigned int)__popc(peers) << 16) | key);
        const unsigned int win = __reduce_max_sync(0xFFFFFFFFu, votes) & 0xFFFFu;
        if (key == win) {
            run_handle(R, E);
            if (run_take(R, C, E)) key = pending_key(R, E);
            else {
                key = 0xFFFFu;
                atomicAdd(p.events_done, run_events(R, C));
                if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
            }
        }
    }
}
