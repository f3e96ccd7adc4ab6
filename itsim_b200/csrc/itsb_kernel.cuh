/*
 * itsb_kernel.cuh -- device side shared by the two translation units that instantiate the engine kernel
 * (itsb_abi.cu: replicas staged in shared memory; itsb_inplace.cu: replicas advanced in place in HBM).  Two
 * translation units so that each kernel's code is laid out on its own: the staged kernel is instruction-fetch
 * bound and sensitive to what else sits around it.
 */
#pragma once
#include <cuda_runtime.h>
#include "itsb_host_common.h"

using namespace itsb;

/* ---- TMA bulk copies (non-tensor form) --------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* ---- kernel parameters ---------------------------------------------------------------------------------------------- */
struct RunParams {
    Model model;            /* device pointers */
    Layout layout;
    unsigned char* hot;
    ArenaNode* arena;
    uint32_t* arena_free;
    itsb_trace_rec* trace;
    itsb_net_event* netlog;
    itsb_connection* connlog;
    ConnEnt* conntab;
    uint64_t n_replicas;
    uint64_t seed;
    double duration;
    unsigned long long* ticket;       /* next replica to claim */
    unsigned long long* events_done;  /* events processed by this launch */
    long long* first_failed;          /* min index of a replica with non-zero status, or LLONG_MAX */
    uint32_t static_bytes;            /* size of the CTA-shared copy of the model tables */
    uint32_t warps_per_cta;
    uint32_t state_in_hbm;            /* 1: replicas too large for shared memory are advanced in place (L1/L2-cached) */
    uint32_t tables_in_hbm;           /* 1: the model tables themselves exceed shared memory: read in place (L1-cached) */
    const Model* model_dev;           /* the Model struct in device memory (pointers to the tables in HBM)             */
    uint32_t pool_stick;              /* k_run_pool: a warp takes the kind the CTA's warps took last while its ring holds
                                         at least this many slots (0: always the fullest ring) -- warps of an SM then
                                         tend to run the same handler, i.e. share fetched instruction lines */
    uint32_t pool_by_program;         /* k_run_pool: WAKE events queued per program (1) or with the other kinds (0) */
    uint32_t lanes_per_warp;          /* k_run_vote: lanes of a warp that take replicas (a batch smaller than the GPU's
                                         threads is spread over all warps instead of filling a few of them)             */
};

struct StaticOffsets { uint32_t total; };

static inline StaticOffsets static_offsets(const itsb_model_desc& d) {
    StaticOffsets o;
    uint32_t off = align_up((uint32_t)sizeof(Model), 16);
    off = align_up(off + 8 * d.n_code, 16);
    off = align_up(off + (uint32_t)sizeof(itsb_entry) * d.n_entries, 16);
    off = align_up(off + (uint32_t)sizeof(itsb_dist) * d.n_dists, 16);
    off = align_up(off + 8 * d.n_consts, 16);
    off = align_up(off + (uint32_t)sizeof(itsb_net) * d.n_nets, 16);
    off = align_up(off + (uint32_t)sizeof(itsb_node) * d.n_nodes, 16);
    off = align_up(off + (uint32_t)sizeof(itsb_filter) * d.n_filters, 16);
    off = align_up(off + (uint32_t)sizeof(itsb_iface) * d.n_ifaces, 16);
    off = align_up(off + (uint32_t)sizeof(itsb_route) * d.n_routes, 16);
    off = align_up(off + 4 * d.n_members, 16);
    o.total = align_up(off, 128);
    return o;
}

__device__ __forceinline__ void cta_copy(void* dst, const void* src, uint32_t bytes) {
    const uint4* s = (const uint4*)src;
    uint4* d = (uint4*)dst;
#pragma unroll 1
    for (uint32_t i = threadIdx.x; i < (bytes + 15) / 16; i += blockDim.x) d[i] = s[i];
}

/* Stages the model tables into shared memory (they are padded to 16 bytes on the device) behind a Model struct whose
 * pointers address the shared copies; returns that struct's address (the out-of-line engine functions reach the
 * tables through it). */
__device__ __forceinline__ const Model* stage_model(const Model& g, unsigned char* sm) {
    Model m = g;
    uint32_t off = ((uint32_t)sizeof(Model) + 15) & ~15u;
    /* same order and element sizes as static_offsets() above */
#define ITSB_STAGE(field, type, count)                                                               \
    cta_copy(sm + off, g.field, (uint32_t)sizeof(type) * (count)); m.field = (const type*)(sm + off); \
    off = (off + (uint32_t)sizeof(type) * (count) + 15) & ~15u;
    ITSB_STAGE(code, u64, g.n_code)
    ITSB_STAGE(entries, itsb_entry, g.n_entries)
    ITSB_STAGE(dists, itsb_dist, g.n_dists)
    ITSB_STAGE(consts, double, g.n_consts)
    ITSB_STAGE(nets, itsb_net, g.n_nets)
    ITSB_STAGE(nodes, itsb_node, g.n_nodes)
    ITSB_STAGE(filters, itsb_filter, g.n_filters)
    ITSB_STAGE(ifaces, itsb_iface, g.n_ifaces)
    ITSB_STAGE(routes, itsb_route, g.n_routes)
    ITSB_STAGE(members, u32, g.n_members)
#undef ITSB_STAGE
    if (threadIdx.x == 0) *(Model*)sm = m;
    __syncthreads();
    return (const Model*)sm;
}

/* ---- the engine kernel ------------------------------------------------------------------------------------------------
 * Two instantiations: replicas staged in shared memory (the compiler then knows every state access is a shared-memory
 * access), and replicas advanced in place in HBM for models too large for that.  One template, because letting a
 * run-time flag pick the pointer made every access a generic-address access and cost 22 % on the staged path. */
template <bool IN_PLACE>
__device__ __forceinline__ void run_warps(const RunParams& p, unsigned char* smem) {
    const Model* m = p.tables_in_hbm ? p.model_dev : stage_model(p.model, smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* staged = smem + p.static_bytes + (size_t)warp * p.layout.hot_bytes;
    uint64_t* bar = (uint64_t*)(smem + p.static_bytes + (IN_PLACE ? 0 : (size_t)p.warps_per_cta * p.layout.hot_bytes)) + warp;
    if (!IN_PLACE) {
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0;
    for (;;) {
        unsigned long long r = 0;
        if (lane == 0) r = atomicAdd(p.ticket, 1ull);
        r = __shfl_sync(0xFFFFFFFFu, r, 0);
        if (r >= p.n_replicas) break;
        unsigned char* hot;
        if (IN_PLACE) {
            hot = p.hot + r * p.layout.hot_bytes;
        } else {
            /* HBM -> shared: one bulk copy of the whole hot block, completion on the warp's mbarrier */
            hot = staged;
            if (lane == 0) {
                mbar_expect_tx(bar, p.layout.hot_bytes);
                bulk_g2s(hot, p.hot + r * p.layout.hot_bytes, p.layout.hot_bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        }
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
        __syncwarp();
        const int32_t status = R.hdr->status;
        if (!IN_PLACE) fence_async_smem();      /* shared -> HBM */
        __syncwarp();
        if (lane == 0) {
            if (!IN_PLACE) {
                bulk_s2g(p.hot + r * p.layout.hot_bytes, hot, p.layout.hot_bytes);
                bulk_wait_all();
            }
            atomicAdd(p.events_done, events);
            if (status != 0) atomicMin(p.first_failed, (long long)r);
        }
        __syncwarp();
    }
}

