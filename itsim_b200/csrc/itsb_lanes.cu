/*
 * itsb_lanes.cu -- a replica per LANE instead of a replica per warp: k_run_vote (the default engine kernel for large
 * batches), and k_run_lanes (the same without the vote: an experiment, built only with -DITSB_EXPERIMENTS).
 *
 * The same engine source at width 1 (what tests/hostemu compiles for the CPU), as device code: every thread advances
 * its own replica in place in HBM, 32 replicas share a warp's instruction stream and diverge wherever their next events
 * differ.  No bucketing by next event yet: this kernel measures the two costs that design has to
 * beat -- divergence without sorting, and per-lane state traffic without an interleaved layout.  Results are
 * bit-identical to k_run's (a replica depends only on (seed, id)).
 */
#include "itsb_vote.cuh"
#if defined(ITSB_EXPERIMENTS)
extern "C" __global__ void __launch_bounds__(256, ITSB_LANES_MIN_CTAS) k_run_lanes(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Model* m = p.tables_in_hbm ? p.model_dev : stage_model(p.model, smem);
    for (;;) {
        const unsigned long long r = atomicAdd(p.ticket, 1ull);
        if (r >= p.n_replicas) break;
        unsigned char* hot = p.hot + r * p.layout.hot_bytes;
        Rep R;
        Binding B;
        B.arena = p.arena + r * p.layout.arena_nodes;
        B.arena_free = p.arena_free + r * p.layout.arena_nodes;
        B.trace = p.trace ? p.trace + r * p.layout.trace_cap : nullptr;
        B.netlog = p.netlog ? p.netlog + r * p.layout.log_cap : nullptr;
        B.connlog = p.connlog ? (itsb_connection*)((unsigned char*)p.connlog + r * conn_log_bytes(p.layout.conn_cap)) : nullptr;
        B.conntab = p.conntab ? p.conntab + r * ((size_t)p.layout.max_sock * 2 * p.layout.conn_window) : nullptr;
        bind_views(R, hot, &p.layout, m, B, p.seed);
        const unsigned long long events = run_replica(R, p.duration);
        atomicAdd(p.events_done, events);
        if (R.hdr->status != 0) atomicMin(p.first_failed, (long long)r);
    }
}

#endif

extern "C" __global__ void __launch_bounds__(256, ITSB_LANES_MIN_CTAS) k_run_vote(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    run_vote(p, smem);
}
