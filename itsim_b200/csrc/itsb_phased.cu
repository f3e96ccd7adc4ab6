/*
 * itsb_phased.cu -- k_run with the warps of a CTA aligned on the two halves of the event loop.
 *
 * k_run is bound by instruction supply (profiles/README.md): its hot code (41 KB) does not fit the SM's 32 KB
 * instruction cache, and twelve warps walking through it independently miss independently.  Here every iteration of the
 * CTA is: (A) each warp pops its next event and does the network work it asks for (transmission start / end, the
 * lane-parallel delivery pass); barrier; (B) each warp that has a process to activate runs the interpreter; barrier.
 * Warps in the same half execute the same few kilobytes at the same time, so a line fetched for one is there for the
 * others.  Event order inside a replica is untouched (a replica is still one warp's); results are bit-identical to
 * k_run's (tests/test_gpu_parity.py runs both).  Its own translation unit, like the other variants.
 */
#include "itsb_kernel.cuh"

#ifndef ITSB_PHASE_POPS
#define ITSB_PHASE_POPS 8
#endif

extern "C" __global__ void __launch_bounds__(384, 1) k_run_phased(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Model* m = p.tables_in_hbm ? p.model_dev : stage_model(p.model, smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* hot = smem + p.static_bytes + (size_t)warp * p.layout.hot_bytes;
    uint64_t* bar = (uint64_t*)(smem + p.static_bytes + (size_t)p.warps_per_cta * p.layout.hot_bytes) + warp;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t phase = 0;
    bool have = false, exhausted = false;
    unsigned long long r = 0;
    Rep R;
    RunCtx C;
    for (;;) {
        /* ---- half A: claim / pop / network ------------------------------------------------------------------------- */
        if (!have && !exhausted) {
            if (lane == 0) r = atomicAdd(p.ticket, 1ull);
            r = __shfl_sync(0xFFFFFFFFu, r, 0);
            if (r >= p.n_replicas) {
                exhausted = true;
            } else {
                if (lane == 0) {
                    mbar_expect_tx(bar, p.layout.hot_bytes);
                    bulk_g2s(hot, p.hot + r * p.layout.hot_bytes, p.layout.hot_bytes, bar);
                }
                mbar_wait(bar, phase);
                phase ^= 1;
                Binding B;
                B.arena = p.arena + r * p.layout.arena_nodes;
                B.arena_free = p.arena_free + r * p.layout.arena_nodes;
                B.trace = p.trace ? p.trace + r * p.layout.trace_cap : nullptr;
                B.netlog = p.netlog ? p.netlog + r * p.layout.log_cap : nullptr;
                B.connlog = p.connlog ? (itsb_connection*)((unsigned char*)p.connlog + r * conn_log_bytes(p.layout.conn_cap)) : nullptr;
                B.conntab = p.conntab ? p.conntab + r * ((size_t)p.layout.max_sock * 2 * p.layout.conn_window) : nullptr;
                bind_views(R, hot, &p.layout, m, B, p.seed);
                run_begin(R, p.duration, C);
                have = true;
            }
        }
        uint32_t runnable = NIL32;
        bool over = false;
        if (have) {
            /* pop (and do the network work of) events until one hands back a process to activate, so that every warp
               reaches the barrier with work for the second half */
#pragma unroll 1
            for (int k = 0; k < ITSB_PHASE_POPS && !over && runnable == NIL32; ++k) over = !run_pop(R, C, runnable);
        }
        __syncthreads();
        /* ---- half B: process activation ------------------------------------------------------------------------------ */
        if (have && runnable != NIL32) run_process(R, runnable);
        if (have && over) {
            const unsigned long long events = run_events(R, C);
            __syncwarp();
            const int32_t status = R.hdr->status;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(p.hot + r * p.layout.hot_bytes, hot, p.layout.hot_bytes);
                bulk_wait_all();
                atomicAdd(p.events_done, events);
                if (status != 0) atomicMin(p.first_failed, (long long)r);
            }
            __syncwarp();
            have = false;
        }
        if (!__syncthreads_or(have || !exhausted)) break;
    }
}
