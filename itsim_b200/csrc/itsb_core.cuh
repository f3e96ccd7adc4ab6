/*
 * itsb_core.cuh -- the per-replica discrete-event engine: one warp advances one replica.
 *
 * What it replaces (reference = pure Python on greensim, /root/reference):
 *   event loop / heap            greensim.Simulator.run/_schedule     (itsim/simulator.py:9)
 *   processes, advance, throw    greensim.Process/advance/interrupt   (itsim/simulator.py:80-83; rows a2)
 *   Signal / Queue waits         greensim.Signal.wait/turn_on         (rows a3); used by sockets and the wake signal
 *   transmission + fan-out       Network.transmit                     (network_simple_overrides.py:273-291; row a16)
 *                                Link._transfer_packet                (itsim/network/link.py:70-86; row a5)
 *   delivery at a node           Node._receive / Node._receive_packet (overrides.py:379-382; node.py:232-243; a7)
 *   sockets, ports               open_socket/bind/get_port_free       (overrides.py:337-369,469-476; rows a8-a9)
 *   variates                     greensim.random + itsim.random       (itsim/random.py:14-29; row a14)
 *
 * Execution model.  All lanes of the warp run this code with warp-uniform control flow; replica state lives in the
 * warp's slice of shared memory.  Scalar bookkeeping is executed redundantly by every lane (same value, same
 * address), so each lane is self-consistent without fences; the lane-parallel phases -- the event-pool minimum
 * (strided scan + redux min-reduction), the delivery fan-out over a link's recipients (ballot / prefix ranks) -- end
 * with a warp sync.  Socket packet queues, remembered-packet lists and packets in flight live in a per-replica
 * arena in HBM (queues grow to thousands of packets while a workstation sleeps).
 *
 * Code layout.  The first profile of this kernel was instruction-fetch bound (60 % of stall samples "no
 * instruction": 100+ KB of SASS, a dozen warps per SM each somewhere else in it).  So the hot loop -- next event,
 * delivery pass, the handful of instructions a daemon executes per packet -- is inlined into the kernel, and
 * everything else (variates, socket open/close, spawn/exit, transmission bookkeeping, spill paths, tracing) is an
 * out-of-line function that takes only the replica header: the header carries the offsets and pointers needed to
 * rebuild a working view (Rep) of the replica on entry.
 *
 * The same source builds with ITSB_W == 1 (one "lane") as a plain C++ translation unit; that build exists only so
 * the unit tests can exercise the interpreter without a GPU (tests/hostemu).  The product library never contains it.
 */
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/itsb.h"
#include "itsb_ir.h"

#if defined(__CUDACC__) && defined(ITSB_LANES)
/* a replica per LANE (itsb_lanes.cu): the width-1 form of the engine -- what the host emulation compiles -- as device
 * code.  Every "warp primitive" is the identity; 32 replicas share a warp's instruction stream and diverge freely. */
#define ITSB_W 1
#define ITSB_FN __device__ __forceinline__
#define ITSB_COLD static __device__ __noinline__
#elif defined(__CUDACC__)
#define ITSB_W 32
#define ITSB_FN __device__ __forceinline__
#define ITSB_COLD static __device__ __noinline__
/* Code placement.  ptxas lays the out-of-line functions of a kernel out behind its body in the order of their
 * mangled names (<length><name>...), and the kernel is instruction-fetch bound (the hot code is larger than the
 * 32 KB instruction cache level, profiles/).  The dozen functions that run per event are therefore given names
 * that sort first, in call order, so the hot code is one contiguous range behind the event loop instead of being
 * scattered over the 200 KB of the shipped-itsim primitives. */
#else
#define ITSB_W 1
#define ITSB_FN static inline
#define ITSB_COLD static __attribute__((noinline))
#endif
#if defined(__CUDACC__)
/* (both device shapes: the lane-per-replica kernels turned out to be bound by instruction supply too) */
#define exec_cold         a0_execute
#define draw_dist         a1_variate
#define log_f64           a1z_ln_f64
#define rng_next          a2_philox_
#define on_tx_begin       a3_txbegin
#define on_tx_end         a4_tx_end_
#define ev_push           a4z_evpush
#define heap_pop          a5_heappop
#define net_send          a6_netsend
#define proc_spawn        a7_spawn__
#define sock_open_on      a8_sk_open
#define sock_close        a9_sk_clos
#define proc_exit         aa_prcexit
#define deliver_exception ab_deliv_x
#endif

namespace itsb {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

static const u16 NIL16 = 0xFFFF;
static const u32 NIL32 = 0xFFFFFFFFu;
static const u64 T_EMPTY = 0x7FF0000000000000ull; /* +inf: empty event slot */
static const u32 NQ_CAP = 128;                    /* now-queue ring entries (power of two) */

/* ---- warp primitives (trivial when ITSB_W == 1) --------------------------------------------------------------- */
#if defined(__CUDACC__) && defined(ITSB_LANES)
ITSB_FN int lane_id() { return 0; }
ITSB_FN u32 w_ballot(bool p) { return p ? 1u : 0u; }
ITSB_FN u32 w_shfl(u32 v, int) { return v; }
ITSB_FN u32 w_min(u32 v) { return v; }
ITSB_FN void w_sync() {}
ITSB_FN u32 w_excl_scan(u32 v, u32* total) { *total = v; return 0; }
ITSB_FN u32 lanemask_lt() { return 0; }
ITSB_FN int popc(u32 v) { return __popc(v); }
ITSB_FN int ffs32(u32 v) { return __ffs(v); }
ITSB_FN u64 dbits(double d) { return (u64)__double_as_longlong(d); }
ITSB_FN double bitsd(u64 b) { return __longlong_as_double((long long)b); }
ITSB_FN u32 mulhi32(u32 a, u32 b) { return __umulhi(a, b); }
ITSB_FN u32 clz32(u32 v) { return (u32)__clz((int)v); }
ITSB_FN u32 popc64(u64 v) { return (u32)__popcll((unsigned long long)v); }
ITSB_FN u32 ctz64(u64 v) { return (u32)(__ffsll((long long)v) - 1); }
#elif defined(__CUDACC__)
ITSB_FN int lane_id() { return (int)(threadIdx.x & 31); }
ITSB_FN u32 w_ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
ITSB_FN u32 w_shfl(u32 v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
ITSB_FN u32 w_min(u32 v) { return __reduce_min_sync(0xFFFFFFFFu, v); }
ITSB_FN void w_sync() { __syncwarp(); }
ITSB_FN u32 w_excl_scan(u32 v, u32* total) {
    u32 x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane_id() >= d) x += y;
    }
    *total = __shfl_sync(0xFFFFFFFFu, x, 31);
    return x - v;
}
ITSB_FN u32 lanemask_lt() { return (1u << lane_id()) - 1u; }
ITSB_FN int popc(u32 v) { return __popc(v); }
ITSB_FN int ffs32(u32 v) { return __ffs(v); }
ITSB_FN u64 dbits(double d) { return (u64)__double_as_longlong(d); }
ITSB_FN double bitsd(u64 b) { return __longlong_as_double((long long)b); }
ITSB_FN u32 mulhi32(u32 a, u32 b) { return __umulhi(a, b); }
ITSB_FN u32 clz32(u32 v) { return (u32)__clz((int)v); }
ITSB_FN u32 popc64(u64 v) { return (u32)__popcll((unsigned long long)v); }
ITSB_FN u32 ctz64(u64 v) { return (u32)(__ffsll((long long)v) - 1); }
#else
ITSB_FN int lane_id() { return 0; }
ITSB_FN u32 w_ballot(bool p) { return p ? 1u : 0u; }
ITSB_FN u32 w_shfl(u32 v, int) { return v; }
ITSB_FN u32 w_min(u32 v) { return v; }
ITSB_FN void w_sync() {}
ITSB_FN u32 w_excl_scan(u32 v, u32* total) { *total = v; return 0; }
ITSB_FN u32 lanemask_lt() { return 0; }
ITSB_FN int popc(u32 v) { return __builtin_popcount(v); }
ITSB_FN int ffs32(u32 v) { return __builtin_ffs((int)v); }
ITSB_FN u64 dbits(double d) { u64 b; memcpy(&b, &d, 8); return b; }
ITSB_FN double bitsd(u64 b) { double d; memcpy(&d, &b, 8); return d; }
ITSB_FN u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }
ITSB_FN u32 clz32(u32 v) { return (u32)__builtin_clz(v); }
ITSB_FN u32 popc64(u64 v) { return (u32)__builtin_popcountll(v); }
ITSB_FN u32 ctz64(u64 v) { return (u32)__builtin_ctzll(v); }
#endif

/* ---- per-replica state ------------------------------------------------------------------------------------------ */

struct Pcb {                /* 48 bytes: one greensim process */
    u16 pc;
    u8 state;
    u8 prog;
    u16 node;
    u16 gen;
    u16 wnext, wprev;       /* waiter-list links */
    u16 wsig;               /* what it waits on: WS_* | index */
    u16 timer_ev;           /* pool slot of its pending advance() wake-up */
    u16 sock;
    u16 exc;                /* last exception delivered */
    u32 list_head, list_tail; /* remembered-packet list (arena nodes) */
    u32 r[4];
    u16 timer_tok;          /* bumped by every advance(); a TIMER event is live only if it carries the current one */
    u8 phase;               /* RECV_FILTER: 0 = before the awake check, 1 = inside recv() */
    u8 tags;                /* greensim tags: the body's and, cascaded, the creator's */
};

enum { PS_FREE = 0, PS_UNSTARTED = 1, PS_RUNNING = 2, PS_TIMER = 3, PS_WAIT = 4, PS_RESUMING = 5 };
enum { WS_NONE = 0, WS_SOCK = 0x4000, WS_NODE = 0x8000, WS_SIG = 0xC000 };

/* -- tables of the shipped-itsim primitives (present only when the model asks for them) -- */
struct PcbX {               /* 40 bytes: extension of a process control block */
    u16 thread;             /* the itsim Thread this computation belongs to (NIL16: a plain greensim process) */
    u16 tnext;              /* next computation of the same thread */
    u32 balker_h;           /* handle of the Queue.join timeout helper, while parked in a timed wait */
    u32 sel_common;         /* select(): the private signal and the two helper processes */
    u32 sel_h1, sel_h2;
    u32 x[4];               /* x[0:2] = f64 delay the computation was started with, x[2:4] = f64 scratch */
    u32 pad;
};

struct Sig {                /* 8 bytes: greensim.Signal */
    u16 head, tail;         /* waiters, FIFO */
    u16 on;
    u16 gen;
};

struct Thr {                /* 20 bytes: itsim Thread */
    u16 task;
    u16 ncomp;
    u16 comp_head, comp_tail;
    u16 dead_sig;
    u16 gen;
    u16 next;               /* next thread of the same process */
    u16 n;
    u16 live;
    u16 pad;
};

struct Task {               /* 24 bytes: itsim Process */
    u16 pid;
    u16 node;
    u32 parent;
    u16 nthreads;
    u16 thread_counter;
    u16 thr_head, thr_tail;
    u16 dead_sig;
    u16 gen;
    u16 live;
    u16 pad;
};

struct Sock {               /* 24 bytes */
    u16 port;
    u16 node;
    u16 next;               /* next socket of the same node */
    u16 wait_head, wait_tail;
    u16 owner;
    u32 q_head, q_tail;     /* FIFO of arena nodes */
    u32 q_len;
};

struct NodeDyn {            /* 16 bytes */
    u16 sock_head;
    u16 port_cursor;        /* next candidate of the unprivileged-port cycle (overrides.py:469) */
    u16 wake_head, wake_tail;
    u32 epoch;              /* bumped by sleep() and wake_up(); stands for _time_awoken (network_simple.py:51-57) */
    u32 awake;
};

struct Pkt {                /* 32 bytes: a packet in flight (an arena node) */
    u32 src_ip, dst_ip;
    u32 ports;              /* src | dst << 16 */
    u32 size;
    u32 tag, f0, f1;
    u32 meta;               /* src node | net << 16 */
};

struct ArenaNode {          /* 32 bytes = one DRAM sector: a queued / remembered packet */
    u32 next;
    u32 src_ip;
    u32 ports;
    u32 size;
    u32 tag, f0, f1;
    u32 pad;
};

struct ConnEnt {            /* 32 bytes: a packet a socket remembers until the other half of its pair shows up */
    u32 key_ip;             /* the peer Location the entry is filed under */
    u32 a_ip;               /* outbound: Remote_IP; inbound: the address the packet was sent to */
    u32 size;
    u32 ports;              /* key port | (outbound: destination port; inbound: the port it was sent to) << 16 */
    u32 t_lo, t_hi;         /* when it was filed */
    u32 pad[2];
};

struct EvNode {             /* arena node used by the now-queue spill chain and the pool overflow list */
    u32 next;
    u32 seq;
    u32 data;
    u32 t_lo, t_hi;
    u32 pad[3];
};

struct SockV {              /* 8 bytes: packets queued "virtually" on a socket (see ND_ASLEEP) */
    u32 count;
    u32 bytes;
};

struct HeapEnt {            /* 16 bytes: one entry of the future-event heap, ordered by (t, seq) */
    u64 t;                  /* timestamp, as its bit pattern (t >= 0: bit order is numeric order) */
    u32 seq;                /* greensim's insertion counter: breaks ties */
    u32 slot;               /* the event's stable slot: ev_data[slot] is its data word, ev_pos[slot] its heap position */
};

struct Layout {             /* byte offsets inside one replica's hot block */
    u32 hot_bytes;
    u32 off_ev_heap, off_ev_pos, off_ev_data, off_ev_free, off_nq;
    u32 off_pcb, off_pcb_free, off_sock, off_sock_free, off_node, off_stats;
    u32 max_ev, max_proc, max_sock, n_nodes;
    u32 arena_nodes, trace_cap;
    /* shipped-itsim tables (all zero when the model does not use them) */
    u32 off_iface, off_pcbx, off_sig, off_sig_free, off_thr, off_thr_free, off_task, off_task_free;
    u32 n_ifaces, max_sig, max_thr, max_task;
    u32 off_tab, table_words;
    u32 log_cap;            /* network_event records per replica (HBM) */
    u32 conn_cap, conn_window, off_sockx, address_tags;   /* connection records (overrides.py:102-186) */
    u32 off_ndig;           /* delivery digests: one u32 per node (see node_digest_compute) */
    u32 off_nmask, off_pmask;   /* per node / per process: which hostname fields the remembered packets carry (64-bit Bloom) */
    u32 off_sockpkt;        /* per socket: the packet at the head of its queue, inline (override-path models) */
    u32 off_sockv;          /* per socket: packets its sleeping owner's loop head will pop and reject (count, bytes) */
};

struct Model {              /* pointers to the static tables (shared-memory copies inside the kernel) */
    const u64* code;
    const itsb_entry* entries;
    const itsb_dist* dists;
    const double* consts;
    const itsb_net* nets;
    const itsb_node* nodes;
    const itsb_initproc* init;
    const itsb_filter* filters;
    const itsb_iface* ifaces;
    const itsb_route* routes;
    const u32* members;
    u32 n_code, n_entries, n_dists, n_consts, n_nets, n_nodes, n_init, n_filters, n_ifaces, n_routes, n_members;
};

/* interpreter scratch registers, valid inside one activation of a process; index = operand source - 0x10 */
enum { VOL_COUNT = 42 };

struct Hdr {                /* first thing in the hot block */
    /* -- persistent state -- */
    double now;
    u64 draws;              /* Philox draw index (oracle/philox.py stream layout) */
    u32 seq;                /* greensim's insertion counter */
    int32_t status;         /* 0, ITSB_ERR_CAP_* or ITSB_EXC_* */
    u64 detail;
    u32 arena_free_top;
    u32 trace_count;
    u32 log_count;          /* network_event records produced */
    u32 conn_count;         /* connection records produced */
    u16 ev_free_top, pcb_free_top, sock_free_top, spare16;
    u32 replica_lo, replica_hi;
    u32 nq_head, nq_tail;   /* now-queue ring cursors (monotone) */
    u32 nq_spill_head, nq_spill_tail, nq_spill_count;
    u32 ev_count;           /* entries in the future-event heap */
    u32 spare32;
    u64 ovf_min_t;
    u32 ovf_min_seq, ovf_head, ovf_count;
    u32 cur_seq;            /* counter value of the event being processed: the order key of the connection records */
    u32 rng_c0, rng_c1;     /* the second half of the Philox block the last even draw computed (= the next draw) */
    u16 sig_free_top, thr_free_top, task_free_top;
    u16 setup_done;         /* the per-replica set-up draws (ITSB_INIT_ADD_DRAWN) have been made */
    u32 hidden_entry_balk, hidden_entry_wait_one;   /* entry indices of greensim's helper programs */
    /* -- binding: rewritten every time the replica is staged (pointers of this launch, not state) -- */
    ArenaNode* arena;
    u32* arena_free;
    itsb_trace_rec* trace;
    itsb_net_event* netlog;
    itsb_connection* connlog;
    ConnEnt* conntab;       /* [socket][direction][window] */
    const Model* model;
    u32 seed_lo, seed_hi;
    Layout L;
    u32 vol[VOL_COUNT];
};

struct Rep {                /* register-resident working view of one replica */
    Hdr* hdr;
    HeapEnt* ev_heap;
    u16* ev_pos;
    u32* ev_data;
    u16* ev_free;
    u32* nq;                /* now-queue ring: (seq, data) pairs */
    Pcb* pcb;
    u16* pcb_free;
    Sock* sock;
    ArenaNode* sockpkt;
    SockV* sockv;
    u16* sock_free;
    NodeDyn* node;
    u32* ndig;
    u64* nmask;
    u64* pmask;
    u64* stats;
    ArenaNode* arena;
    u32* arena_free;
    itsb_trace_rec* trace;
    Model m;
    u32 max_ev, trace_cap, max_sig;
    u32* iface_addr;        /* shipped-itsim tables */
    PcbX* pcbx;
    Sig* sig;
    u16* sig_free;
    Thr* thr;
    u16* thr_free;
    Task* task;
    u16* task_free;
    u32* tab;               /* model table (TLOAD / TSTORE) */
};

/* Rebuilds the working view from the header alone (used on entry by every out-of-line function; unused fields are
 * dead code there). */
ITSB_FN void rebind(Rep& R, Hdr* H) {
    unsigned char* hot = (unsigned char*)H;
    R.hdr = H;
    R.ev_heap = (HeapEnt*)(hot + H->L.off_ev_heap);
    R.ev_pos = (u16*)(hot + H->L.off_ev_pos);
    R.ev_data = (u32*)(hot + H->L.off_ev_data);
    R.ev_free = (u16*)(hot + H->L.off_ev_free);
    R.nq = (u32*)(hot + H->L.off_nq);
    R.pcb = (Pcb*)(hot + H->L.off_pcb);
    R.pcb_free = (u16*)(hot + H->L.off_pcb_free);
    R.sock = (Sock*)(hot + H->L.off_sock);
    R.sockpkt = (ArenaNode*)(hot + H->L.off_sockpkt);
    R.sockv = (SockV*)(hot + H->L.off_sockv);
    R.sock_free = (u16*)(hot + H->L.off_sock_free);
    R.node = (NodeDyn*)(hot + H->L.off_node);
    R.ndig = (u32*)(hot + H->L.off_ndig);
    R.nmask = (u64*)(hot + H->L.off_nmask);
    R.pmask = (u64*)(hot + H->L.off_pmask);
    R.stats = (u64*)(hot + H->L.off_stats);
    R.arena = H->arena;
    R.arena_free = H->arena_free;
    R.trace = H->trace;
    R.m = *H->model;
    R.max_ev = H->L.max_ev;
    R.trace_cap = H->L.trace_cap;
    R.max_sig = H->L.max_sig;
    R.iface_addr = (u32*)(hot + H->L.off_iface);
    R.pcbx = (PcbX*)(hot + H->L.off_pcbx);
    R.sig = (Sig*)(hot + H->L.off_sig);
    R.sig_free = (u16*)(hot + H->L.off_sig_free);
    R.thr = (Thr*)(hot + H->L.off_thr);
    R.thr_free = (u16*)(hot + H->L.off_thr_free);
    R.task = (Task*)(hot + H->L.off_task);
    R.task_free = (u16*)(hot + H->L.off_task_free);
    R.tab = (u32*)(hot + H->L.off_tab);
}

/* Stages the launch's pointers into the header, then builds the view. */
struct Binding {            /* one replica's regions outside the hot block, for this launch */
    ArenaNode* arena;
    u32* arena_free;
    itsb_trace_rec* trace;
    itsb_net_event* netlog;
    itsb_connection* connlog;
    ConnEnt* conntab;
};

ITSB_FN void bind_views(Rep& R, unsigned char* hot, const Layout* L, const Model* model, const Binding& B, u64 seed) {
    Hdr* H = (Hdr*)hot;
    H->L = *L;
    H->model = model;
    H->arena = B.arena;
    H->arena_free = B.arena_free;
    H->trace = B.trace;
    H->netlog = B.netlog;
    H->connlog = B.connlog;
    H->conntab = B.conntab;
    H->seed_lo = (u32)seed;
    H->seed_hi = (u32)(seed >> 32);
    w_sync();
    rebind(R, H);
}

ITSB_COLD void fail(Hdr* H, int status, u64 detail) {
    if (H->status == 0) {
        H->status = status;
        H->detail = detail;
    }
}

/* delivery digests (defined with the sockets, below) */
ITSB_COLD void node_digest_refresh(Hdr* H, u32 node);

/* ---- event data word: kind[0:4] exc[4:12] proc[12:25] gen[25:32]; packet events: kind[0:4] arena-index[4:32] ---- */
static const u32 PROC_MASK = 0x1FFF;   /* up to 8192 process slots per replica */
static const u32 GEN_MASK = 0x7F;      /* slot generation: tells a finished process from the one that reuses its slot */
ITSB_FN u32 ev_pack(u32 kind, u32 idx, u32 gen, u32 exc) {
    return kind | ((exc & 0xFF) << 4) | ((idx & PROC_MASK) << 12) | ((gen & GEN_MASK) << 25);
}
ITSB_FN u32 ev_kind(u32 d) { return d & 0xF; }
ITSB_FN u32 ev_exc(u32 d) { return (d >> 4) & 0xFF; }
ITSB_FN u32 ev_idx(u32 d) { return (d >> 12) & PROC_MASK; }
ITSB_FN u32 ev_gen(u32 d) { return (d >> 25) & GEN_MASK; }
ITSB_FN u32 ev_pack_pkt(u32 kind, u32 node_index) { return kind | (node_index << 4); }
ITSB_FN u32 ev_pkt(u32 d) { return d >> 4; }

/* ---- Philox4x32-10 stream (same definition as oracle/philox.py) --------------------------------------------------- */
ITSB_FN void philox_block(u32 seed_lo, u32 seed_hi, u32 rep_lo, u32 rep_hi, u64 block, u32 out[4]) {
    u32 c0 = (u32)block, c1 = (u32)(block >> 32), c2 = seed_hi, c3 = rep_hi;
    u32 k0 = seed_lo, k1 = rep_lo;
    for (int rnd = 0; rnd < 10; ++rnd) {
        if (rnd) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        u32 hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        u32 hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        u32 n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Draw d of the replica's stream: the 64 bits (a, b) of its half of Philox block d >> 1. */
ITSB_COLD u64 rng_next(Hdr* H) {
    const u64 d = H->draws;
    H->draws = d + 1;
    if (d & 1) return ((u64)H->rng_c0 << 32) | H->rng_c1;   /* draws are sequential: draw d - 1 computed this block */
    u32 w[4];
    philox_block(H->seed_lo, H->seed_hi, H->replica_lo, H->replica_hi, d >> 1, w);
    H->rng_c0 = w[2]; H->rng_c1 = w[3];
    return ((u64)w[0] << 32) | w[1];
}

ITSB_FN double rng_random(Hdr* H) {  /* CPython's 53-bit construction */
    const u64 ab = rng_next(H);
    const u32 a = (u32)(ab >> 32), b = (u32)ab;
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

/* math.log, one copy of its code for both variates that need it -- and ONE ALGORITHM on every side of the parity
 * check.  CPython's expovariate / normalvariate call the platform libm's log, whose last bit is not specified (glibc,
 * CUDA: both <= 1 ulp, not the same ulp); a 1-ulp difference flips int(min(max(x + 68, 0), 576)) or the
 * Kinderman-Monahan acceptance test once in ~1e13 draws and desynchronises the replica for good.  So the logarithm is
 * written out here with IEEE +, -, *, / only (this file is compiled with -fmad=false / -ffp-contract=off, no table):
 * the FreeBSD msun / fdlibm e_log.c algorithm (reduce x = 2^k (1 + f), sqrt(2)/2 <= 1 + f < sqrt(2); s = f / (2 + f);
 * log(1 + f) = f - (f^2/2 - s (f^2/2 + R(s^2))) with a degree-14 minimax R; error < 1 ulp).  oracle/philox.py log_f64
 * is the same sequence of operations in Python; tests/test_philox.py checks the two bit for bit. */
ITSB_COLD double log_f64(double x) {
    const double ln2_hi = bitsd(0x3FE62E42FEE00000ull), ln2_lo = bitsd(0x3DEA39EF35793C76ull);
    const double Lg1 = bitsd(0x3FE5555555555593ull), Lg2 = bitsd(0x3FD999999997FA04ull),
                 Lg3 = bitsd(0x3FD2492494229359ull), Lg4 = bitsd(0x3FCC71C51D8E78AFull),
                 Lg5 = bitsd(0x3FC7466496CB03DEull), Lg6 = bitsd(0x3FC39A09D078C69Full),
                 Lg7 = bitsd(0x3FC2F112DF3E5244ull);
    u64 b = dbits(x);
    int32_t hx = (int32_t)(b >> 32);
    u32 lx = (u32)b;
    int32_t k = 0;
    if (hx < 0x00100000) {                          /* x < 2^-1022 */
        if (((hx & 0x7FFFFFFF) | lx) == 0) return bitsd(0xFFF0000000000000ull);   /* log(+-0) = -inf */
        if (hx < 0) return bitsd(0x7FF8000000000000ull);                           /* log(-#) = NaN */
        k -= 54;
        x *= 18014398509481984.0;                   /* subnormal: scale up by 2^54 */
        b = dbits(x);
        hx = (int32_t)(b >> 32);
        lx = (u32)b;
    }
    if (hx >= 0x7FF00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000FFFFF;
    int32_t i = (hx + 0x95F64) & 0x100000;
    x = bitsd(((u64)(u32)(hx | (i ^ 0x3FF00000)) << 32) | lx);   /* x or x / 2, in [sqrt(2)/2, sqrt(2)) */
    k += i >> 20;
    const double f = x - 1.0;
    const double dk = (double)k;
    if ((0x000FFFFF & (2 + hx)) < 3) {              /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * ln2_hi + dk * ln2_lo;
        }
        const double r = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - r;
        return dk * ln2_hi - ((r - dk * ln2_lo) - f);
    }
    const double s = f / (2.0 + f);
    const double z = s * s;
    i = hx - 0x6147A;
    const double w = z * z;
    const int32_t j = 0x6B851 - hx;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    const double r = t2 + t1;
    if (i > 0) {
        const double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + r));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + r) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - r);
    return dk * ln2_hi - ((s * (f - r) - dk * ln2_lo) - f);
}

/* One value of a generator chain: kind -> linear -> bounded -> project_int (itsim/random.py:29). */
ITSB_COLD double draw_dist(Hdr* H, const itsb_dist* Dp) {
    const itsb_dist D = *Dp;
    double x;
    switch (D.kind) {
    case ITSB_DIST_CONSTANT:
        x = D.p0;
        break;
    case ITSB_DIST_UNIFORM:  /* random.uniform: a + (b - a) * random() */
        x = D.p0 + (D.p1 - D.p0) * rng_random(H);
        break;
    case ITSB_DIST_EXPO:     /* random.expovariate(lambd): -log(1 - random()) / lambd */
        x = -log_f64(1.0 - rng_random(H)) / D.p0;
        break;
    case ITSB_DIST_NORMAL: { /* random.normalvariate: Kinderman-Monahan ratio of uniforms */
        double z;
        for (;;) {
            double u1 = rng_random(H);
            double u2 = 1.0 - rng_random(H);
            z = 1.7155277699214135 * (u1 - 0.5) / u2;
            double zz = z * z / 4.0;
            if (zz <= 1.0 - u2) break;     /* -log(u2) >= 1 - u2 (exact: u2 = 1 - k 2^-53): accepted without the log */
            if (zz <= -log_f64(u2)) break;
        }
        x = D.p0 + z * D.p1;
        break;
    }
    case ITSB_DIST_CHOICE: { /* random.choice -> _randbelow_with_getrandbits(n) */
        const u32 n = (u32)D.p0;
        const u32 k = 32 - clz32(n);
        u32 r;
        for (;;) {
            r = (u32)(rng_next(H) >> 32) >> (32 - k);
            if (r < n) break;
        }
        x = (double)r;
        break;
    }
    default:
        x = 0.0;
    }
    if (D.flags & ITSB_DF_LINEAR) x = D.slope * x + D.shift;
    if ((D.flags & ITSB_DF_LOWER) && x < D.lower) x = D.lower;
    if ((D.flags & ITSB_DF_UPPER) && x > D.upper) x = D.upper;
    if (D.flags & ITSB_DF_INT) x = (double)(long long)x;
    return x;
}

/* ---- telemetry ------------------------------------------------------------------------------------------------------ */
/* (sockv_add / sockv_take below follow the same rule as the counters) */
/* A counter update.  In the lane-per-replica kernel the counters live in HBM and a read-modify-write makes the lane wait
 * for the load (round 2 profile: ~5 % of all stall samples on these lines); a replica's counters are touched by its own
 * lane only, so the update goes out as a reduction (RED.ADD at the L2, nothing to wait for) and the few reads (the
 * run's event count, STAT_FIRST) bypass the L1, which a reduction does not update. */
#if defined(__CUDACC__) && defined(ITSB_LANES)
ITSB_FN void stat_add(Rep& R, u32 slot, u64 v) { atomicAdd((unsigned long long*)&R.stats[slot], (unsigned long long)v); }
ITSB_FN u64 stat_get(Rep& R, u32 slot) { return __ldcg((const unsigned long long*)&R.stats[slot]); }
/* a packet queued virtually on socket s (ND_ASLEEP), and the loop head collecting what was queued */
ITSB_FN void sockv_add(Rep& R, u32 s, u32 bytes) { atomicAdd(&R.sockv[s].count, 1u); atomicAdd(&R.sockv[s].bytes, bytes); }
ITSB_FN SockV sockv_take(Rep& R, u32 s) {
    SockV V;
    V.count = __ldcg(&R.sockv[s].count);
    if (V.count != 0) { V.bytes = __ldcg(&R.sockv[s].bytes); R.sockv[s].count = 0; R.sockv[s].bytes = 0; } else V.bytes = 0;
    return V;
}
#else
ITSB_FN void stat_add(Rep& R, u32 slot, u64 v) { R.stats[slot] += v; }
ITSB_FN u64 stat_get(Rep& R, u32 slot) { return R.stats[slot]; }
ITSB_FN void sockv_add(Rep& R, u32 s, u32 bytes) { R.sockv[s].count += 1; R.sockv[s].bytes += bytes; }
ITSB_FN SockV sockv_take(Rep& R, u32 s) {
    const SockV V = R.sockv[s];
    if (V.count != 0) { R.sockv[s].count = 0; R.sockv[s].bytes = 0; }
    return V;
}
#endif

ITSB_COLD void trace_put(Hdr* H, u32 kind, u32 prog, u32 node, u32 aux) {
    u32 n = H->trace_count;
    if (n < H->L.trace_cap && lane_id() == 0) {
        itsb_trace_rec rec;
        rec.t = H->now; rec.kind = (u8)kind; rec.prog = (u8)prog; rec.node = (u16)node; rec.aux = aux;
        H->trace[n] = rec;
    }
    H->trace_count = n + 1;
}

/* one `network_event` record (itsim/schemas/items.py:83-131): a socket opened (type 1) or closed (type 2) */
ITSB_COLD void net_event_put(Hdr* H, u32 type, u32 node, u32 port, u32 src_ip, u32 prog, int pid, u32 protocol) {
    if (H->netlog == nullptr) return;
    { Rep Rs; Rs.stats = (u64*)((unsigned char*)H + H->L.off_stats); stat_add(Rs, ITSB_TM_USER0 + 7, 1); }   /* records produced (user slot 7) */
    const u32 n = H->log_count;
    if (n < H->L.log_cap && lane_id() == 0) {
        itsb_net_event E;
        E.t = H->now; E.node = (u16)node; E.port = (u16)port; E.type = (u8)type; E.protocol = (u8)protocol;
        E.pid = (int16_t)pid; E.src_ip = src_ip; E.dst_ip = 0; E.dst_port = 0; E.prog = (u16)(prog & 0xFFFF);
        E.tags = (u16)(prog >> 16); E.reserved = 0;
        H->netlog[n] = E;
    }
    H->log_count = n + 1;
}

/* one simulated event of the given kind (SURVEY.md section 8-d definition) */
ITSB_FN void count_event(Rep& R, u32 kind, u32 prog, u32 node, u32 aux) {
    stat_add(R, ITSB_TM_KIND0 + kind, 1);
    if (R.trace) trace_put(R.hdr, kind, prog, node, aux);
}

/* transit-delay histogram: bin = clamp(floor(log2(delay / 1us)), 0, 63) via the exponent field */
ITSB_FN u32 latency_bin(double delay) {
    double scaled = delay * 1.0e6;
    if (!(scaled >= 1.0)) return 0;
    u32 e = (u32)((dbits(scaled) >> 52) & 0x7FF) - 1023u;
    return e > (ITSB_LAT_BINS - 1) ? (ITSB_LAT_BINS - 1) : e;
}

/* ---- arena (HBM) -------------------------------------------------------------------------------------------------- */
ITSB_FN u32 arena_alloc(Rep& R) {
    u32 top = R.hdr->arena_free_top;
    if (top == 0) { fail(R.hdr, ITSB_ERR_CAP_ARENA, 0); return NIL32; }
    top -= 1;
    R.hdr->arena_free_top = top;
    return R.arena_free[top];
}

ITSB_FN void arena_release(Rep& R, u32 idx) {
    u32 top = R.hdr->arena_free_top;
    R.arena_free[top] = idx;   /* every lane stores the same word: each lane stays self-consistent */
    R.hdr->arena_free_top = top + 1;
}

/* ---- events ---------------------------------------------------------------------------------------------------------
 * greensim keeps one heap keyed (t, insertion counter).  The same total order is kept here by two structures:
 *
 *  now-queue   every event scheduled with zero delay (resume, throw, add, the delivery fan-out, TX_BEGIN) has
 *              t == now and a counter larger than anything scheduled before it, so these events are already sorted:
 *              a FIFO ring in shared memory, O(1) push and pop.  They are ~95 % of all pops in network_simple.
 *  future heap events with t > now (advance wake-ups, TX_END): a 4-ary min-heap on (t, seq), below.
 *
 * Neither is bounded in the reference, so both continue into arena nodes in HBM when full (a workstation that wakes
 * with a thousand queued packets emits its answers in one instant): the ring spills to a FIFO chain, the heap to an
 * unordered overflow list with its own cached minimum.
 */
ITSB_COLD void nq_spill_push(Hdr* H, u32 seq, u32 data) {
    Rep R; rebind(R, H);
    u32 n = arena_alloc(R);
    if (n == NIL32) return;
    EvNode E;
    E.next = NIL32; E.seq = seq; E.data = data; E.t_lo = 0; E.t_hi = 0; E.pad[0] = 0; E.pad[1] = 0; E.pad[2] = 0;
    *(EvNode*)&R.arena[n] = E;
    if (H->nq_spill_count == 0) H->nq_spill_head = n; else ((EvNode*)&R.arena[H->nq_spill_tail])->next = n;
    H->nq_spill_tail = n;
    H->nq_spill_count += 1;
}

ITSB_COLD void ovf_push(Hdr* H, u64 tb, u32 seq, u32 data) {
    Rep R; rebind(R, H);
    u32 n = arena_alloc(R);
    if (n == NIL32) return;
    EvNode E;
    E.next = H->ovf_count ? H->ovf_head : NIL32; E.seq = seq; E.data = data;
    E.t_lo = (u32)tb; E.t_hi = (u32)(tb >> 32); E.pad[0] = 0; E.pad[1] = 0; E.pad[2] = 0;
    *(EvNode*)&R.arena[n] = E;
    H->ovf_head = n;
    if (H->ovf_count == 0 || tb < H->ovf_min_t || (tb == H->ovf_min_t && seq < H->ovf_min_seq)) {
        H->ovf_min_t = tb; H->ovf_min_seq = seq;
    }
    H->ovf_count += 1;
}

/* Moves spilled entries back into the ring once it has drained. */
ITSB_COLD void nq_refill(Hdr* H) {
    Rep R; rebind(R, H);
    while (H->nq_spill_count != 0 && H->nq_tail - H->nq_head < NQ_CAP) {
        u32 n = H->nq_spill_head;
        const EvNode E = *(const EvNode*)&R.arena[n];
        u32 i = H->nq_tail & (NQ_CAP - 1);
        R.nq[2 * i] = E.seq;
        R.nq[2 * i + 1] = E.data;
        H->nq_tail += 1;
        H->nq_spill_head = E.next;
        H->nq_spill_count -= 1;
        arena_release(R, n);
    }
}

/* Removes the minimum of the overflow list (rare: only after the pool filled up) and recomputes its cached minimum. */
ITSB_COLD u32 ovf_take_min(Hdr* H) {
    Rep R; rebind(R, H);
    u32 prev = NIL32, n = H->ovf_head, data = 0;
    const u64 want_t = H->ovf_min_t;
    const u32 want_seq = H->ovf_min_seq;
    u64 nt = T_EMPTY; u32 ns = NIL32;
    bool taken = false;
    u32 left = H->ovf_count;
    while (left != 0) {
        const EvNode E = *(const EvNode*)&R.arena[n];
        const u64 tb = ((u64)E.t_hi << 32) | E.t_lo;
        if (!taken && tb == want_t && E.seq == want_seq) {
            data = E.data;
            if (prev == NIL32) H->ovf_head = E.next; else ((EvNode*)&R.arena[prev])->next = E.next;
            arena_release(R, n);
            taken = true;
        } else {
            if (tb < nt || (tb == nt && E.seq < ns)) { nt = tb; ns = E.seq; }
            prev = n;
        }
        n = E.next;
        left -= 1;
    }
    H->ovf_count -= 1;
    H->ovf_min_t = nt; H->ovf_min_seq = ns;
    return data;
}

/* ---- the future-event heap --------------------------------------------------------------------------------------------
 * A 4-ary min-heap on (t, seq) of 16-byte entries; entry 0 sits at byte 48 of a 64-byte line, so the four children of
 * any entry (4i + 1 .. 4i + 4) are one aligned 64-byte line: a pop costs one memory round per level (three or four for
 * the ~100 sleepers of network_simple) where the unordered pool it replaces was scanned from end to end after every
 * removal (round 1: 22 % of k_run_vote's instructions, 16 dependent rounds of loads per scan).  An event keeps a stable
 * slot (ev_data / ev_pos / the free stack) so that an interrupted advance() can still cancel its wake-up exactly
 * (greensim: "on Interrupt cancel the wake-up"): ev_pos[slot] is kept current by every move.  The peek is entry 0. */
ITSB_FN bool key_less(u64 ta, u32 sa, u64 tb, u32 sb) { return ta < tb || (ta == tb && sa < sb); }

ITSB_FN void heap_sift_up(Rep& R, u32 i, const HeapEnt E) {
    HeapEnt* hp = R.ev_heap;
#if ITSB_W == 1
    /* a lane per replica waits on memory, not on instructions: the ancestors' addresses do not depend on their
       contents, so the first four levels are loaded together (one round instead of up to four) */
    if (i > 0) {
        const u32 a0 = (i - 1) >> 2, a1 = a0 ? (a0 - 1) >> 2 : 0, a2 = a1 ? (a1 - 1) >> 2 : 0, a3 = a2 ? (a2 - 1) >> 2 : 0;
        const HeapEnt A0 = hp[a0], A1 = hp[a1], A2 = hp[a2], A3 = hp[a3];
        const HeapEnt anc[4] = {A0, A1, A2, A3};
        const u32 idx[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i == 0 || !key_less(E.t, E.seq, anc[k].t, anc[k].seq)) break;
            hp[i] = anc[k]; R.ev_pos[anc[k].slot] = (u16)i;
            i = idx[k];
        }
    }
#endif
    while (i > 0) {
        const u32 par = (i - 1) >> 2;
        const HeapEnt A = hp[par];
        if (!key_less(E.t, E.seq, A.t, A.seq)) break;
        hp[i] = A; R.ev_pos[A.slot] = (u16)i;
        i = par;
    }
    hp[i] = E; R.ev_pos[E.slot] = (u16)i;
}

/* places E at position i of a heap of n entries, moving it down while a child precedes it */
ITSB_FN void heap_sift_down(Rep& R, u32 i, const HeapEnt E, u32 n) {
    HeapEnt* hp = R.ev_heap;
    for (;;) {
        const u32 c = 4 * i + 1;
        if (c >= n) break;
        /* the four children: one aligned 64-byte line (absent ones read as +inf) */
        HeapEnt C0 = hp[c], C1, C2, C3;
        C1.t = T_EMPTY; C1.seq = NIL32; C1.slot = 0; C2 = C1; C3 = C1;
        if (c + 1 < n) C1 = hp[c + 1];
        if (c + 2 < n) C2 = hp[c + 2];
        if (c + 3 < n) C3 = hp[c + 3];
        u32 m = c; HeapEnt M = C0;
        if (key_less(C1.t, C1.seq, M.t, M.seq)) { M = C1; m = c + 1; }
        if (key_less(C2.t, C2.seq, M.t, M.seq)) { M = C2; m = c + 2; }
        if (key_less(C3.t, C3.seq, M.t, M.seq)) { M = C3; m = c + 3; }
        if (!key_less(M.t, M.seq, E.t, E.seq)) break;
        hp[i] = M; R.ev_pos[M.slot] = (u16)i;
        i = m;
    }
    hp[i] = E; R.ev_pos[E.slot] = (u16)i;
}

/* the time of the earliest future event (+inf when there is none) */
ITSB_FN u64 heap_min_t(Hdr* H) {
    return H->ev_count ? ((const HeapEnt*)((unsigned char*)H + H->L.off_ev_heap))[0].t : T_EMPTY;
}

/* Removes the event in `slot` wherever it sits in the heap (an interrupted advance() cancels its wake-up). */
ITSB_COLD void ev_release_cold(Hdr* H, u32 slot) {
    Rep R; rebind(R, H);
    const u32 i = R.ev_pos[slot];
    const u32 n = H->ev_count - 1;
    H->ev_count = n;
    u32 top = H->ev_free_top;
    R.ev_free[top] = (u16)slot;
    H->ev_free_top = (u16)(top + 1);
    if (i == n) return;
    const HeapEnt L = R.ev_heap[n];
    if (i > 0) {
        const HeapEnt A = R.ev_heap[(i - 1) >> 2];
        if (key_less(L.t, L.seq, A.t, A.seq)) { heap_sift_up(R, i, L); return; }
    }
    heap_sift_down(R, i, L, n);
}
ITSB_FN void ev_release(Rep& R, u32 slot) { ev_release_cold(R.hdr, slot); }

/* Schedules (t, seq = counter++).  Returns the event's slot (so an advance() can be cancelled) or -1.  One copy of the
 * code for all callers: the kernel is bound by instruction fetch, not by instruction count. */
ITSB_COLD int ev_push(Hdr* H, u64 tb, u32 data) {
    const u32 seq = H->seq;
    H->seq = seq + 1;
    unsigned char* hot = (unsigned char*)H;
    if (tb == dbits(H->now)) {
        if (H->nq_spill_count == 0 && H->nq_tail - H->nq_head < NQ_CAP) {
            u32* nq = (u32*)(hot + H->L.off_nq);
            u32 i = H->nq_tail & (NQ_CAP - 1);
            nq[2 * i] = seq;
            nq[2 * i + 1] = data;
            H->nq_tail += 1;
        } else {
            nq_spill_push(H, seq, data);
        }
        return -1;
    }
    u32 top = H->ev_free_top;
    if (top == 0) {   /* every slot taken: unordered overflow list in the arena */
        ovf_push(H, tb, seq, data);
        return -1;
    }
    top -= 1;
    H->ev_free_top = (u16)top;
    const u32 slot = ((u16*)(hot + H->L.off_ev_free))[top];
    ((u32*)(hot + H->L.off_ev_data))[slot] = data;
    Rep R; rebind(R, H);
    HeapEnt E; E.t = tb; E.seq = seq; E.slot = slot;
    const u32 n = H->ev_count;
    H->ev_count = n + 1;
    heap_sift_up(R, n, E);
    return (int)slot;
}
ITSB_FN int ev_push_inl(Rep& R, double t, u32 data) { return ev_push(R.hdr, dbits(t), data); }

/* Removes the heap's minimum; returns its data word. */
ITSB_COLD u32 heap_pop(Hdr* H) {
    Rep R; rebind(R, H);
    const u32 slot = R.ev_heap[0].slot;
    const u32 d = R.ev_data[slot];
    u32 top = H->ev_free_top;
    R.ev_free[top] = (u16)slot;
    H->ev_free_top = (u16)(top + 1);
    const u32 n = H->ev_count - 1;
    H->ev_count = n;
    if (n != 0) {
        const HeapEnt L = R.ev_heap[n];
        heap_sift_down(R, 0, L, n);
    }
    return d;
}

enum { SRC_NONE_LEFT = 0, SRC_NQ = 1, SRC_POOL = 2, SRC_OVF = 3 };

/* Key of the next event in (t, seq) order and where it sits; nothing is removed yet. */
ITSB_FN int ev_peek(Rep& R, u64& tbits, u32& seq) {
    Hdr* H = R.hdr;
    if (H->nq_head == H->nq_tail && H->nq_spill_count != 0) nq_refill(H);
    int src = SRC_NONE_LEFT;
    tbits = T_EMPTY; seq = NIL32;
    if (H->ev_count != 0) { const HeapEnt E = R.ev_heap[0]; src = SRC_POOL; tbits = E.t; seq = E.seq; }
    if (H->ovf_count != 0 && (H->ovf_min_t < tbits || (H->ovf_min_t == tbits && H->ovf_min_seq < seq))) {
        src = SRC_OVF; tbits = H->ovf_min_t; seq = H->ovf_min_seq;
    }
    if (H->nq_head != H->nq_tail) {
        const u64 nb = dbits(H->now);
        const u32 ns = R.nq[2 * (H->nq_head & (NQ_CAP - 1))];
        if (nb < tbits || (nb == tbits && ns < seq)) { src = SRC_NQ; tbits = nb; seq = ns; }
    }
    return src;
}

/* Removes the event ev_peek chose; returns its data word. */
ITSB_FN u32 ev_take(Rep& R, int src) {
    Hdr* H = R.hdr;
    if (src == SRC_NQ) {
        u32 d = R.nq[2 * (H->nq_head & (NQ_CAP - 1)) + 1];
        H->nq_head += 1;
        return d;
    }
    if (src == SRC_POOL) return heap_pop(H);
    return ovf_take_min(H);
}

/* ---- delivery digests ----------------------------------------------------------------------------------------------
 * What the delivery pass (on_deliver) wants to know about a recipient, in one word per node, kept beside the tables it
 * summarises: a broadcast reaches all 51 nodes of network_simple, and finding out what each does with it by walking
 * node -> socket list -> waiting process -> its instruction -> its filter (-> the packets it remembers) is four or more
 * dependent loads per recipient on state that lives in HBM for the lane-per-replica kernel (round 1: 6.6 DRAM sectors
 * per counted event, 85 % of stall samples waiting on those loads).  The digests of a network's nodes are contiguous:
 * 51 x 4 bytes = 7 sectors.
 *
 *   port[0:16]  filter[16:19]  flags[19:24]  socket[24:31]  ND_QUEUED[31]
 *   ND_SINGLE   the node has exactly one socket, bound to `port` (no socket at all: ND_SINGLE with port 0, which is
 *               never bound) -- a packet for any other port finds nothing: dropped without looking further
 *   ND_RETIRE   the socket bound to `port` has exactly one waiter, parked inside recv() of a RECV_FILTER loop head
 *               (phase 1) using filter `filter`, with the node still in the wake epoch the loop remembered, the queue
 *               empty.  A packet its filter rejects can then be retired on the spot (see on_deliver) from this word
 *               and the static tables alone.  ND_LISTED: that process remembers at least one packet; nmask[node] then
 *               holds a 64-bit Bloom filter (bit = hostname field & 63) of the packets it remembers, a copy of the
 *               process's pmask: a packet whose bit is clear is certainly not "listed"
 *   ND_ASLEEP   the socket bound to `port` has no waiter and no queued packet, and its owner is parked on the node's
 *               wake signal at the head of a RECV_FILTER loop (phase 0: `with ws.awake()`) using filter `filter`, with no
 *               exception handler of its own; ND_LISTED / nmask as above.  Every packet that reaches the socket from now
 *               on will be popped by that loop head in ONE activation when the workstation wakes; until a packet
 *               arrives that the filter lets through, nothing can change what the filter sees, so a packet it rejects
 *               now is certain to be popped, counted and rejected then.  Such a packet is queued *virtually*: counted
 *               in sockv[socket] (count, bytes) and credited by the loop head when it resumes -- no arena node, no
 *               chain of dependent loads to drain a backlog of hundreds of packets on wake-up (round-2 profile of
 *               k_run_pool: 8 % of all stall samples).  The first packet the filter lets through is queued for real and
 *               withdraws the statement -- unless the assembler found the loop behind the head *stable* (ND_STABLE:
 *               whatever the program does with a packet it is handed, it does without blocking and comes back to this
 *               head): then the statement stands, ND_QUEUED records that a real packet is queued, and only packets whose
 *               fate depends on the remembered packets (which a packet queued ahead might change) go to the tables
 *   socket      with any flag: the index of the socket bound to `port` (ND_NOSOCK: not stated), so a packet that
 *               is accepted finds its socket without walking the node's list
 *
 * A digest is a set of STATEMENTS, each of which must be true whenever it is present; an absent statement (flags 0) only
 * sends the pass back to the tables.  The points that can falsify a statement update the word: a socket opened /
 * closed, a process parking on or leaving a socket's wait list, a delivery waking a socket's waiters, the node
 * sleeping or waking (epoch), and -- for pmask -- the three list instructions.  tests/hostemu builds with
 * ITSB_CHECK_DIGEST verify every statement against the tables before every use and fail the replica otherwise. */
enum { ND_STABLE = 1u << 19, ND_ASLEEP = 1u << 20, ND_RETIRE = 1u << 21, ND_LISTED = 1u << 22, ND_SINGLE = 1u << 23,
       ND_QUEUED = 1u << 31, ND_NOSOCK = 0x7Fu, ND_MAX_FILTER = 7u };
ITSB_FN u32 nd_port(u32 d) { return d & 0xFFFFu; }
ITSB_FN u32 nd_filter(u32 d) { return (d >> 16) & 7u; }
ITSB_FN u32 nd_sock(u32 d) { return (d >> 24) & 0x7Fu; }
ITSB_FN u32 nd_sock_field(u32 s) { return (s < ND_NOSOCK ? s : (u32)ND_NOSOCK) << 24; }
/* what is left of a digest when nothing can be retired (or queued virtually) on its socket any more */
ITSB_FN u32 nd_without_retire(u32 d) { return (d & ND_SINGLE) ? (d & (0xFFFFu | ND_SINGLE | 0x7F000000u)) : 0u; }

/* Is process w parked, asleep, at the head of a RECV_FILTER loop on socket s, so that packets reaching s are certain to
 * be popped by that loop head in one activation when the workstation wakes?  (ND_ASLEEP, below.)  *stable: the
 * assembler found the loop behind the head stable (itsim_b200/ir.py _stable_loop): packets handed to the program are
 * handled without blocking and the program comes back to this head, so the statement survives queued packets. */
ITSB_FN bool nd_owner_asleep(Rep& R, u32 node, u32 s, u32 w, u32* filter, bool* stable) {
    const Sock& S = R.sock[s];
    if (w == NIL16 || S.wait_head != NIL16) return false;
    const Pcb& P = R.pcb[w];
    if (P.state != PS_WAIT || P.wsig != (u16)(WS_NODE | node) || P.phase != 0 || P.sock != s) return false;
    const u64 ins = R.m.code[P.pc];
    *filter = (u32)(ins >> 16) & 0xFF;
    *stable = ((u32)(ins >> 24) & 1u) != 0;
    if (S.q_len != 0 && !*stable) return false;
    /* no handler of its own: an exception ends the process through CLOSE; EXIT, which drops the queue either way */
    return (u32)(ins & 0xFF) == OP_RECV_FILTER && (u32)(ins >> 48) == 0 && *filter <= ND_MAX_FILTER;
}

/* the statements recomputed from the tables (socket closed, exception delivered: the rare points) */
ITSB_COLD void node_digest_refresh(Hdr* H, u32 node) {
    Rep R; rebind(R, H);
    if (H->L.max_sig != 0) { R.ndig[node] = 0; return; }   /* shipped-itsim sockets wait on signals: m_deliver */
    const NodeDyn& N = R.node[node];
    u32 n_socks = 0, d = 0, first_port = 0, first_sock = 0;
    bool have = false;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (u16 s = N.sock_head; s != NIL16; s = R.sock[s].next) {
        const Sock& S = R.sock[s];
        if (n_socks == 0) { first_port = S.port; first_sock = s; }
        n_socks += 1;
        const u16 w = S.wait_head;
        u32 fa = 0;
        bool stable = false;
        if (!have && nd_owner_asleep(R, node, s, S.owner, &fa, &stable)) {
            have = true;
            d = S.port | (fa << 16) | ND_ASLEEP | (stable ? (u32)ND_STABLE : 0u) | (S.q_len != 0 ? (u32)ND_QUEUED : 0u) |
                (R.pcb[S.owner].list_head != NIL32 ? ND_LISTED : 0u) | nd_sock_field(s);
            R.nmask[node] = R.pmask[S.owner];
            continue;
        }
        if (have || w == NIL16 || S.q_len != 0) continue;
        const Pcb& P = R.pcb[w];
        if (P.wnext != NIL16 || P.phase != 1) continue;
        const u64 ins = R.m.code[P.pc];
        const u32 wop = (u32)(ins & 0xFF), filter = (u32)(ins >> 16) & 0xFF;
        if ((wop != OP_RECV_FILTER && wop != OP_RECV_FILTER_LOG) || filter > ND_MAX_FILTER) continue;
        if (N.epoch != P.r[(u32)(ins >> 8) & 3]) continue;
        have = true;
        d = S.port | (filter << 16) | ND_RETIRE | (P.list_head != NIL32 ? ND_LISTED : 0u) | nd_sock_field(s);
        R.nmask[node] = R.pmask[w];
    }
    if (n_socks == 0) d = ND_SINGLE | nd_sock_field(ND_NOSOCK);                    /* port 0: matches nothing */
    else if (n_socks == 1) d = have ? (d | ND_SINGLE) : (first_port | ND_SINGLE | nd_sock_field(first_sock));
    R.ndig[node] = d;
}

/* a process parked inside recv() of a RECV_FILTER loop head on socket s (queue empty): everything is at hand */
ITSB_FN void node_digest_parked(Rep& R, u32 node, u32 s, u32 p, u32 filter, u32 epoch_reg) {
    if (R.hdr->L.max_sig != 0) return;
    const NodeDyn& N = R.node[node];
    const Sock& S = R.sock[s];
    const Pcb& P = R.pcb[p];
    const u32 old = R.ndig[node];
    const u32 single = (N.sock_head == s && S.next == NIL16) ? (u32)ND_SINGLE : 0u;
    if (S.wait_head == p && P.wnext == NIL16 && N.epoch == P.r[epoch_reg & 3] && filter <= ND_MAX_FILTER) {
        R.ndig[node] = S.port | (filter << 16) | ND_RETIRE | (P.list_head != NIL32 ? ND_LISTED : 0u) | single | nd_sock_field(s);
        if (P.list_head != NIL32) R.nmask[node] = R.pmask[p];
    } else if (nd_port(old) == S.port) {
        R.ndig[node] = single ? (S.port | single | nd_sock_field(s)) : 0u;   /* a second waiter, or a stale epoch */
    }
}

/* a process parked on socket s by a plain recv(): whatever was said about retiring on that socket no longer holds */
ITSB_FN void node_digest_waits(Rep& R, u32 node, u32 s) {
    if (R.hdr->L.max_sig != 0) return;
    const u32 old = R.ndig[node];
    if ((old & (ND_RETIRE | ND_ASLEEP)) && nd_port(old) == R.sock[s].port) R.ndig[node] = nd_without_retire(old);
}

/* a delivery woke every waiter of the socket bound to `port` */
ITSB_FN void node_digest_woken(Rep& R, u32 node, u32 port) {
    const u32 old = R.ndig[node];
    if ((old & (ND_RETIRE | ND_ASLEEP)) && nd_port(old) == port) R.ndig[node] = nd_without_retire(old);
}

/* a packet was queued for real on the socket bound to `port`: nothing may be queued virtually behind it */
ITSB_FN void node_digest_enqueued(Rep& R, u32 node, u32 port) {
    const u32 old = R.ndig[node];
    if ((old & ND_ASLEEP) && nd_port(old) == port)
        R.ndig[node] = (old & ND_STABLE) ? (old | ND_QUEUED) : nd_without_retire(old);
}

/* process p parked on its node's wake signal at the head of a RECV_FILTER loop (phase 0) */
ITSB_FN void node_digest_asleep(Rep& R, u32 node, u32 p) {
    if (R.hdr->L.max_sig != 0) return;
    const Pcb& P = R.pcb[p];
    const u32 s = P.sock;
    u32 filter = 0;
    bool stable = false;
    if (s == NIL16 || !nd_owner_asleep(R, node, s, p, &filter, &stable) || R.sock[s].owner != p) return;
    const NodeDyn& N = R.node[node];
    const Sock& S = R.sock[s];
    const u32 single = (N.sock_head == s && S.next == NIL16) ? (u32)ND_SINGLE : 0u;
    R.ndig[node] = S.port | (filter << 16) | ND_ASLEEP | (stable ? (u32)ND_STABLE : 0u) | (S.q_len != 0 ? (u32)ND_QUEUED : 0u) |
                   (P.list_head != NIL32 ? ND_LISTED : 0u) | single | nd_sock_field(s);
    if (P.list_head != NIL32) R.nmask[node] = R.pmask[p];
}

/* the node slept or woke: its epoch moved on, so a loop head parked in recv() remembers a stale one */
ITSB_FN void node_digest_epoch(Rep& R, u32 node) {
    const u32 old = R.ndig[node];
    if (old & (ND_RETIRE | ND_ASLEEP)) R.ndig[node] = nd_without_retire(old);
}

/* socket s was just put at the head of the node's socket list */
ITSB_FN void node_digest_opened(Rep& R, u32 node, u32 s) {
    if (R.hdr->L.max_sig != 0) return;
    const Sock& S = R.sock[s];
    const u32 old = R.ndig[node];
    if (S.next == NIL16) R.ndig[node] = S.port | ND_SINGLE | nd_sock_field(s);     /* the node's only socket */
    else R.ndig[node] = (old & (ND_RETIRE | ND_ASLEEP)) ? (old & ~(u32)ND_SINGLE) : 0u;   /* what was said about another socket stands */
}

#if defined(ITSB_CHECK_DIGEST)
/* every statement of the stored digest, verified against the tables (test builds only) */
static bool node_digest_valid(Rep& R, u32 node) {
    const u32 d = R.ndig[node];
    const NodeDyn& N = R.node[node];
    if (R.hdr->L.max_sig != 0) return d == 0;
    if (d & ND_SINGLE) {
        if (nd_port(d) == 0) { if (N.sock_head != NIL16) return false; }
        else if (N.sock_head == NIL16 || R.sock[N.sock_head].next != NIL16 || R.sock[N.sock_head].port != nd_port(d)) return false;
    }
    if ((d & ND_RETIRE) && (d & ND_ASLEEP)) return false;
    if ((d & (ND_SINGLE | ND_RETIRE | ND_ASLEEP)) && nd_port(d) != 0 && nd_sock(d) != ND_NOSOCK) {
        const Sock& S = R.sock[nd_sock(d)];
        if (S.port != nd_port(d) || S.node != node) return false;
        bool linked = false;
        for (u16 s = N.sock_head; s != NIL16; s = R.sock[s].next) linked |= (s == nd_sock(d));
        if (!linked) return false;
    }
    if (d & ND_RETIRE) {
        u16 s = N.sock_head;
        while (s != NIL16 && R.sock[s].port != nd_port(d)) s = R.sock[s].next;
        if (s == NIL16) return false;
        const Sock& S = R.sock[s];
        if (S.wait_head == NIL16 || S.q_len != 0) return false;
        const Pcb& P = R.pcb[S.wait_head];
        const u64 ins = R.m.code[P.pc];
        const u32 wop = (u32)(ins & 0xFF);
        if (P.wnext != NIL16 || P.phase != 1 || (wop != OP_RECV_FILTER && wop != OP_RECV_FILTER_LOG)) return false;
        if (((u32)(ins >> 16) & 0xFF) != nd_filter(d) || N.epoch != P.r[(u32)(ins >> 8) & 3]) return false;
        if (((d & ND_LISTED) != 0) != (P.list_head != NIL32)) return false;
        u64 want = 0;
        for (u32 n = P.list_head; n != NIL32; n = R.arena[n].next) want |= 1ull << (R.arena[n].f0 & 63);
        if (R.pmask[S.wait_head] != want) return false;
        if ((d & ND_LISTED) && R.nmask[node] != want) return false;
    } else if (d & ND_ASLEEP) {
        if (d & ND_RETIRE) return false;
        const u32 s = nd_sock(d);
        if (s == ND_NOSOCK || R.sock[s].port != nd_port(d) || R.sock[s].node != node) return false;
        u32 f = 0;
        bool stable = false;
        const u32 w = R.sock[s].owner;
        if (!nd_owner_asleep(R, node, s, w, &f, &stable) || f != nd_filter(d) || stable != ((d & ND_STABLE) != 0)) return false;
        if (((d & ND_QUEUED) != 0) != (R.sock[s].q_len != 0)) return false;
        const Pcb& P = R.pcb[w];
        if (((d & ND_LISTED) != 0) != (P.list_head != NIL32)) return false;
        u64 want = 0;
        for (u32 n = P.list_head; n != NIL32; n = R.arena[n].next) want |= 1ull << (R.arena[n].f0 & 63);
        if (R.pmask[w] != want || ((d & ND_LISTED) && R.nmask[node] != want)) return false;
    } else if (d & (ND_LISTED | ND_QUEUED)) return false;
    return true;
}
#endif

/* ---- processes ----------------------------------------------------------------------------------------------------------- */
ITSB_FN u32 handle_of(Rep& R, u32 p) { return p | ((u32)R.pcb[p].gen << 16); }

/* Creates a process in PS_UNSTARTED and schedules its first switch (greensim add / add_in). */
ITSB_COLD int proc_spawn(Hdr* H, u32 entry, u32 node, const u32* regs, double delay) {
    Rep R; rebind(R, H);
    u32 top = H->pcb_free_top;
    if (top == 0) { fail(H, ITSB_ERR_CAP_PROCS, 0); return -1; }
    top -= 1;
    H->pcb_free_top = (u16)top;
    const u32 p = R.pcb_free[top];
    Pcb& P = R.pcb[p];
    const itsb_entry E = R.m.entries[entry];
    const u32 r0 = regs[0], r1 = regs[1], r2 = regs[2], r3 = regs[3];
    P.pc = (u16)E.pc;
    P.state = PS_UNSTARTED;
    P.prog = (u8)E.prog;
    P.tags = (u8)(E.tags | H->vol[SRC_TAGS - 0x10]);   /* the creator is the process running now (none at set-up) */
    P.node = (u16)node;
    P.wnext = NIL16; P.wprev = NIL16; P.wsig = WS_NONE; P.timer_ev = NIL16; P.sock = NIL16; P.exc = 0;
    P.list_head = NIL32; P.list_tail = NIL32; P.phase = 0;
    P.r[0] = r0; P.r[1] = r1; P.r[2] = r2; P.r[3] = r3;
    if (H->L.max_sig != 0) {
        PcbX& X = R.pcbx[p];
        X.thread = NIL16; X.tnext = NIL16; X.balker_h = NIL32; X.sel_common = NIL16; X.sel_h1 = NIL32; X.sel_h2 = NIL32;
        X.x[0] = 0; X.x[1] = 0; X.x[2] = 0; X.x[3] = 0; X.pad = 0;
    }
    ev_push_inl(R, H->now + delay, ev_pack(ITSB_EV_PROC_START, p, P.gen, 0));
    return (int)p;
}

ITSB_FN void list_free_all(Rep& R, Pcb& P) {
    u32 n = P.list_head;
    while (n != NIL32) {
        u32 nx = R.arena[n].next;
        arena_release(R, n);
        n = nx;
    }
    P.list_head = NIL32; P.list_tail = NIL32;
}

ITSB_COLD void proc_exit(Hdr* H, u32 p) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    if (P.list_head != NIL32) { list_free_all(R, P); R.pmask[p] = 0; }
    P.state = PS_FREE;
    P.gen = (u16)((P.gen + 1) & GEN_MASK);
    u32 top = H->pcb_free_top;
    R.pcb_free[top] = (u16)p;
    H->pcb_free_top = (u16)(top + 1);
}

/* ---- waiter lists (greensim Queue: FIFO by join order) --------------------------------------------------------------- */
ITSB_FN void waiters_append(Rep& R, u16& head, u16& tail, u32 p, u16 wsig) {
    Pcb& P = R.pcb[p];
    P.wnext = NIL16;
    P.wprev = tail;
    P.wsig = wsig;
    if (tail == NIL16) head = (u16)p; else R.pcb[tail].wnext = (u16)p;
    tail = (u16)p;
}

ITSB_FN void waiters_unlink(Rep& R, u16& head, u16& tail, u32 p) {
    Pcb& P = R.pcb[p];
    if (P.wprev == NIL16) head = P.wnext; else R.pcb[P.wprev].wnext = P.wnext;
    if (P.wnext == NIL16) tail = P.wprev; else R.pcb[P.wnext].wprev = P.wprev;
    P.wnext = NIL16; P.wprev = NIL16; P.wsig = WS_NONE;
}

/* Signal.turn_on: resume every waiter in order (each resume = _schedule(0, switch)). */
ITSB_FN void waiters_resume_all(Rep& R, u16& head, u16& tail) {
    u16 p = head;
    while (p != NIL16) {
        Pcb& P = R.pcb[p];
        u16 nx = P.wnext;
        P.wnext = NIL16; P.wprev = NIL16; P.wsig = WS_NONE;
        P.state = PS_RESUMING;
        ev_push_inl(R, R.hdr->now, ev_pack(ITSB_EV_WAKE, p, P.gen, 0));
        p = nx;
    }
    head = NIL16; tail = NIL16;
}

/* ws.wake_up(): signal on, every waiter resumed in join order, then the wake time changes (network_simple.py:51-53) */
ITSB_COLD void node_wake(Hdr* H, u32 node) {
    Rep R; rebind(R, H);
    NodeDyn& N = R.node[node];
    N.awake = 1;
    waiters_resume_all(R, N.wake_head, N.wake_tail);
    N.epoch += 1;
    node_digest_epoch(R, node);
}

/* the ordered fallback of the delivery pass: one socket's waiters through the spilling push */
ITSB_COLD void sock_resume_waiters(Hdr* H, u32 s) {
    Rep R; rebind(R, H);
    Sock& S = R.sock[s];
    waiters_resume_all(R, S.wait_head, S.wait_tail);
    node_digest_woken(R, S.node, S.port);
}

/* ---- sockets ------------------------------------------------------------------------------------------------------------ */
ITSB_FN int sock_find(Rep& R, u32 node, u32 dst_ip, u32 dst_port) {
    const itsb_node N = R.m.nodes[node];
    if (dst_ip != N.addr && dst_ip != R.m.nets[N.net].bcast) return -1;
    u16 s = R.node[node].sock_head;
    while (s != NIL16) {
        if (R.sock[s].port == dst_port) return (int)s;
        s = R.sock[s].next;
    }
    return -1;
}

ITSB_FN bool port_in_use(Rep& R, u32 node, u32 port) {
    u16 s = R.node[node].sock_head;
    while (s != NIL16) {
        if (R.sock[s].port == port) return true;
        s = R.sock[s].next;
    }
    return false;
}

/* open_socket (overrides.py:337-369): port 0 takes the next free port of the persistent 1024..65535 cycle. */
ITSB_COLD int sock_open_on(Hdr* H, u32 p, u32 node, u32 port, bool own);
ITSB_FN int sock_open(Hdr* H, u32 p, u32 port) {
    return sock_open_on(H, p, ((Pcb*)((unsigned char*)H + H->L.off_pcb))[p].node, port, true);
}

/* `own`: the socket becomes the process's current one (send / recv / close act on it).  A socket opened on another
 * attachment of the machine (malware_simple.py:98-99) is only ever named by FORWARD. */
ITSB_COLD int sock_open_on(Hdr* H, u32 p, u32 node, u32 port, bool own) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    NodeDyn& N = R.node[node];
    if (port == 0) {
        u32 c = N.port_cursor;
        for (;;) {
            u32 cand = c;
            c = (c == 65535u) ? 1024u : c + 1u;
            if (!port_in_use(R, node, cand)) { port = cand; break; }
        }
        N.port_cursor = (u16)c;
    } else if (port_in_use(R, node, port)) {
        fail(H, ITSB_EXC_PORT_IN_USE, port);
        return -1;
    }
    u32 top = H->sock_free_top;
    if (top == 0) { fail(H, ITSB_ERR_CAP_SOCKETS, port); return -1; }
    top -= 1;
    H->sock_free_top = (u16)top;
    u32 s = R.sock_free[top];
    Sock& S = R.sock[s];
    S.port = (u16)port; S.node = (u16)node; S.owner = (u16)p;
    S.wait_head = NIL16; S.wait_tail = NIL16;
    S.q_head = NIL32; S.q_tail = NIL32; S.q_len = 0;
    R.sockv[s].count = 0; R.sockv[s].bytes = 0;
    S.next = N.sock_head;
    N.sock_head = (u16)s;
    if (own) P.sock = (u16)s;
    node_digest_opened(R, node, s);
    net_event_put(H, 1, node, port, R.m.nodes[node].addr, P.prog | ((u32)P.tags << 16), 0, 1);
    return (int)s;
}

ITSB_COLD void sock_close(Hdr* H, u32 p) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    u32 s = P.sock;
    if (s == NIL16) return;
    Sock& S = R.sock[s];
    NodeDyn& N = R.node[S.node];
    net_event_put(H, 2, S.node, S.port, R.m.nodes[S.node].addr, P.prog | ((u32)P.tags << 16), 0, 1);
    if (N.sock_head == s) N.sock_head = S.next;
    else {
        u16 q = N.sock_head;
        while (q != NIL16 && R.sock[q].next != s) q = R.sock[q].next;
        if (q != NIL16) R.sock[q].next = S.next;
    }
    u32 n = S.q_head;      /* drop queued packets */
    while (n != NIL32) {
        u32 nx = R.arena[n].next;
        arena_release(R, n);
        n = nx;
    }
    S.q_head = NIL32; S.q_tail = NIL32; S.q_len = 0;
    R.sockv[s].count = 0; R.sockv[s].bytes = 0;      /* packets queued virtually are dropped with the rest */
    S.port = 0;
    u32 top = H->sock_free_top;
    R.sock_free[top] = (u16)s;
    H->sock_free_top = (u16)(top + 1);
    P.sock = NIL16;
    node_digest_refresh(H, S.node);
}

/* A socket's packet queue (Socket._packet_queue, a FIFO).  In override-path models the packet at the HEAD sits inline,
 * in the socket's own 32-byte slot of the hot block (sockpkt), and only the packets behind it are arena nodes: a daemon
 * that is awake and waiting has an empty queue, so a packet it accepts goes into the slot and comes out of it again
 * without an arena node being allocated, written, read back and released (round 2 profile of k_run_vote: those four
 * steps were two of the four dependent DRAM round trips of an accepted packet).  q_len counts the slot and the nodes.
 * Shipped-itsim models (Link mode, itsb_machine.cuh) keep every queued packet in arena nodes. */
ITSB_FN bool sock_head_inline(Rep& R) { return R.max_sig == 0; }

/* the packet recv() would return next (caller checked q_len != 0) */
ITSB_FN ArenaNode sock_front(Rep& R, Sock& S) {
    return sock_head_inline(R) ? R.sockpkt[(u32)(&S - R.sock)] : R.arena[S.q_head];
}

/* Pops the head of the process's socket queue (caller checked q_len != 0). */
ITSB_FN ArenaNode sock_pop(Rep& R, Sock& S) {
    ArenaNode A;
    if (sock_head_inline(R)) {
        const u32 s = (u32)(&S - R.sock);
        A = R.sockpkt[s];
        S.q_len -= 1;
        if (S.q_len != 0) {                       /* the next one moves up into the slot */
            const u32 n = S.q_head;
            ArenaNode B = R.arena[n];
            S.q_head = B.next;
            if (B.next == NIL32) S.q_tail = NIL32;
#if defined(__CUDACC__) && defined(ITSB_LANES)
            /* a workstation that wakes up drains a backlog of hundreds of packets through this chain of dependent
               loads (round 2 profile of k_run_pool: 5 % of all stall samples): the node after this one is fetched now,
               while the loop head still filters this packet */
            else asm volatile("prefetch.global.L2 [%0];" ::"l"(R.arena + B.next));
#endif
            B.next = NIL32;
            R.sockpkt[s] = B;
            arena_release(R, n);
        }
    } else {
        const u32 n = S.q_head;
        A = R.arena[n];
        S.q_head = A.next;
        if (A.next == NIL32) S.q_tail = NIL32;
        S.q_len -= 1;
        arena_release(R, n);
    }
    stat_add(R, ITSB_TM_PACKETS_RECEIVED, 1);
    stat_add(R, ITSB_TM_BYTES_RECEIVED, A.size);
    return A;
}

/* ---- transmission ---------------------------------------------------------------------------------------------------------- */
ITSB_FN Pkt* pkt_at(Rep& R, u32 k) { return (Pkt*)&R.arena[k]; }

/* index of the node of network NET that owns address ip, or NIL32 (first match in link order).  Addresses handed out by
 * Network.link are consecutive (cidr.hosts(), overrides.py:207-213), so the guess ip - addr(first) is checked first. */
ITSB_FN u32 net_member(Rep& R, const itsb_net& NET, u32 ip) {
    const u32 guess = ip - R.m.nodes[NET.first_node].addr;
    u32 hit = NIL32;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (u32 i = (u32)lane_id(); i < NET.n_nodes; i += ITSB_W) {
        if (R.m.nodes[NET.first_node + i].addr == ip) { hit = i; if (ITSB_W == 1) break; }
        if (ITSB_W == 1 && i == 0 && guess < NET.n_nodes && guess != 0 && R.m.nodes[NET.first_node + guess].addr == ip) {
            /* only a shortcut when no earlier node has the same address: addresses on a network are unique
               (AddressInUse, overrides.py:236-237), so the guess is the first match */
            hit = guess; break;
        }
    }
    return w_min(hit);
}

/* Socket.send -> Node._send_to_network -> Network.transmit (overrides.py:122-125,371-377,273-291): the packet
 * becomes a transmission process started with sim.add, i.e. one TX_BEGIN event at the current time. */
ITSB_COLD void net_send(Hdr* H, u32 src_node, u32 src_port, u32 dst_ip, u32 dst_port, u32 size, u32 tag, u32 f0,
                        u32 f1) {
    Rep R; rebind(R, H);
    const itsb_node N = R.m.nodes[src_node];
    const itsb_net NET = R.m.nets[N.net];
    if (dst_ip != NET.bcast) {   /* must be a member of the network (CannotForward otherwise, overrides.py:293-297) */
        if (net_member(R, NET, dst_ip) == NIL32) { fail(H, ITSB_EXC_NO_SUCH_ADDRESS, dst_ip); return; }
    }
    u32 k = arena_alloc(R);
    if (k == NIL32) return;
    Pkt K;
    K.src_ip = N.addr; K.dst_ip = dst_ip;
    K.ports = src_port | (dst_port << 16);
    K.size = size; K.tag = tag; K.f0 = f0; K.f1 = f1;
    K.meta = src_node | (N.net << 16);
    *pkt_at(R, k) = K;
    stat_add(R, ITSB_TM_PACKETS_SENT, 1);
    stat_add(R, ITSB_TM_BYTES_SENT, size);
    ev_push_inl(R, H->now, ev_pack_pkt(ITSB_EV_TX_BEGIN, k));
}

/* transmission(): advance(next(latency) + len(packet) / next(bandwidth))  (overrides.py:287-288) */
ITSB_COLD void on_tx_begin(Hdr* H, u32 k) {
    Rep R; rebind(R, H);
    const Pkt K = *pkt_at(R, k);
    const itsb_net NET = R.m.nets[K.meta >> 16];
    count_event(R, ITSB_EV_TX_BEGIN, 9, K.meta & 0xFFFF, K.size);
    double lat = draw_dist(H, &R.m.dists[NET.lat_dist]);
    const itsb_dist* BW = &R.m.dists[NET.bw_dist];
    double bw = (BW->kind == ITSB_DIST_CONSTANT && (BW->flags & (ITSB_DF_LINEAR | ITSB_DF_UPPER | ITSB_DF_INT)) == 0)
                    ? ((BW->flags & ITSB_DF_LOWER) && BW->p0 < BW->lower ? BW->lower : BW->p0)   /* constant(c): no draw */
                    : draw_dist(H, BW);
    double delay = lat + (double)K.size / bw;
    if (delay < 0.0) { fail(H, ITSB_EXC_NEGATIVE_DELAY, k); return; }
    stat_add(R, ITSB_TM_LAT0 + latency_bin(delay), 1);
    ev_push_inl(R, H->now + delay, ev_pack_pkt(ITSB_EV_TX_END, k));
}

/* for node in receivers: sim.add(node._receive, packet)  -- one queue entry stands for the R consecutive pushes.
 * Returns true when that entry would be the very next event anyway (nothing queued at this instant, no future event
 * at t == now): the caller then runs the delivery pass directly instead of pushing and popping it. */
ITSB_COLD bool on_tx_end(Hdr* H, u32 k) {
    Rep R; rebind(R, H);
    const Pkt K = *pkt_at(R, k);
    const itsb_net NET = R.m.nets[K.meta >> 16];
    u32 count = (K.dst_ip == NET.bcast) ? NET.n_nodes : 1u;
    count_event(R, ITSB_EV_TX_END, 9, K.meta & 0xFFFF, K.size);
    if (H->nq_head == H->nq_tail && H->nq_spill_count == 0) {
        const u64 nb = dbits(H->now);
        if (heap_min_t(H) > nb && (H->ovf_count == 0 || H->ovf_min_t > nb)) {
            H->seq += count;
            return true;
        }
    }
    ev_push_inl(R, H->now, ev_pack_pkt(ITSB_EV_DELIVER, k));
    H->seq += count - 1;
    return false;
}

ITSB_COLD void m_deliver(Hdr* H, u32 k);   /* Link-mode delivery, itsb_machine.cuh */

/* does the process remember a packet whose hostname field is f0 (what LIST_TAKE would find)? */
ITSB_FN bool list_holds(Rep& R, const Pcb& P, u32 f0) {
    u32 n = P.list_head;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    while (n != NIL32) {
        if (R.arena[n].f0 == f0) return true;
        n = R.arena[n].next;
    }
    return false;
}

ITSB_FN bool filter_passes(Rep& R, const itsb_filter F, u32 tag, u32 f0, u32 host, const Pcb& P) {
    const u32 bit = 1u << (tag & 31);
    if ((F.tags_always & bit) != 0 || ((F.tags_if_host & bit) != 0 && f0 == host) ||
        ((F.tags_if_list & bit) != 0 && P.list_head != NIL32))
        return true;
    return (F.tags_if_listed & bit) != 0 && P.list_head != NIL32 && list_holds(R, P, f0);
}

/* trace records of one delivery pass: one DELIVER record per recipient, written by its lane */
ITSB_COLD void trace_deliveries(Hdr* H, u32 first_slot, u32 node, bool valid, u32 aux, u32 n_valid) {
    if (valid) {
        u32 n = H->trace_count + first_slot;
        if (n < H->L.trace_cap) {
            itsb_trace_rec rec;
            rec.t = H->now; rec.kind = ITSB_EV_DELIVER; rec.prog = 10; rec.node = (u16)node; rec.aux = aux;
            H->trace[n] = rec;
        }
    }
    w_sync();
    H->trace_count += n_valid;
}

/* Node._receive for every recipient, in link order (overrides.py:379-382), lane-parallel.
 *
 * Retiring filtered wake-ups in the delivery pass.  A recipient whose only waiter sits in a RECV_FILTER loop, awake,
 * with an empty queue, and whose filter rejects this packet would be woken, pop the packet, reject it and block
 * again on the same socket: no state changes but the counters.  When nothing else is pending at this instant
 * (now-queue empty, no future event at t == now) that outcome cannot be influenced by anything in between, so the
 * lane retires the wake-up on the spot: no queue node, no event, same counters, same counter values consumed.  Off
 * while tracing, so a trace always shows the reference's interleaving event by event. */
ITSB_COLD void conn_retired(Hdr* H, u32 mask, u32 s_lane, u32 seq_lane, u32 k);

ITSB_FN void on_deliver(Rep& R, u32 k) {
    Hdr* H = R.hdr;
    const Pkt K = *pkt_at(R, k);
    const itsb_net NET = R.m.nets[K.meta >> 16];
    if (NET.mode == ITSB_NET_LINK) { m_deliver(H, k); return; }
    const bool bcast = (K.dst_ip == NET.bcast);
    u32 first = NET.first_node, count = NET.n_nodes;
    if (!bcast) {
        first = NET.first_node + net_member(R, NET, K.dst_ip);
        count = 1;
    }
    const u32 dst_port = K.ports >> 16;
    const double now = H->now;
    bool fast_ok = false;
    if (R.trace == nullptr && R.m.n_filters != 0) {   /* (a model that keeps connection records has no RECV_FILTER) */
        if (H->nq_head == H->nq_tail && H->nq_spill_count == 0) {
            const u64 nb = dbits(now);
            fast_ok = heap_min_t(H) > nb && (H->ovf_count == 0 || H->ovf_min_t > nb);
        }
    }
    /* what a recipient's digest may decide on its own: "no socket on that port" always; "retired" when retiring is on
       and the retired recv()'s correlation is not wanted (a model that keeps connection records does it per socket) */
    const bool digest_retires = fast_ok && H->conntab == nullptr;
    const u32 tag_bit = 1u << (K.tag & 31);
    const bool inl = sock_head_inline(R);
    enum { RC_TABLES = 0, RC_DROPPED = 1, RC_RETIRED = 2, RC_VIRTUAL = 3 };

    /* a recipient's digest: nobody listens / retired on the spot / queued virtually / look at the tables (s = its
       socket when the digest states it) */
    const bool virtual_ok = R.trace == nullptr && H->conntab == nullptr;
    auto classify = [&](u32 node, int& s) -> u32 {
        const u32 dg = R.ndig[node];
#if defined(ITSB_CHECK_DIGEST)
        if (!node_digest_valid(R, node)) { fail(H, ITSB_ERR_BAD_OPCODE, 0xD16E57ull << 32 | node); return RC_DROPPED; }
#endif
        if (nd_port(dg) != dst_port) return (dg & ND_SINGLE) ? (u32)RC_DROPPED : (u32)RC_TABLES;   /* its only socket is elsewhere */
        if ((dg & (ND_SINGLE | ND_RETIRE | ND_ASLEEP)) && nd_sock(dg) != ND_NOSOCK) s = (int)nd_sock(dg);
        const bool retires = digest_retires && (dg & ND_RETIRE);
        const bool sleeps = virtual_ok && (dg & ND_ASLEEP) && nd_sock(dg) != ND_NOSOCK;   /* (queued on a socket it can name) */
        if (retires || sleeps) {
            const itsb_filter F = R.m.filters[nd_filter(dg)];
            if ((F.tags_always & tag_bit) != 0 || ((F.tags_if_host & tag_bit) != 0 && K.f0 == R.m.nodes[node].host))
                return RC_TABLES;
            /* handed over while it remembers a packet (if_list), or one with this hostname field (if_listed): the
               Bloom bit says "certainly not" for the second; anything else is for the tables to decide */
            const bool list_class = ((F.tags_if_list | F.tags_if_listed) & tag_bit) != 0;
            if (dg & ND_LISTED) {
                if (F.tags_if_list & tag_bit) return RC_TABLES;
                if ((F.tags_if_listed & tag_bit) && ((R.nmask[node] >> (K.f0 & 63)) & 1ull) != 0) return RC_TABLES;
            }
            if (retires) return RC_RETIRED;
            /* asleep: what the filter makes of a tag that depends on the remembered packets can still change if a
               packet queued ahead is handed to the program first */
            if (list_class && (dg & ND_QUEUED)) return RC_TABLES;
            return RC_VIRTUAL;
        }
        return RC_TABLES;
    };

#if ITSB_W == 1
    u32 q_valid = 0, q_found = 0, q_retired = 0;       /* the pass's counters, flushed once */
#endif
    /* One pass through the tables: the recipients first + base + lane (one per lane; one in all at width 1).  `known`:
       the caller classified them already and these are the ones left for the tables. */
    auto pass = [&](u32 base, bool known) -> bool {
        u32 i = base + (u32)lane_id();
        bool valid = i < count;
        u32 node = first + i;
        int s = -1;
        bool found = false, retire = false, queued_virtually = false;
        if (valid) {
            u32 rc = RC_TABLES;
            if (known) {     /* classified by the caller and left for the tables: only the socket hint is read again */
                const u32 dg = R.ndig[node];
                if (nd_port(dg) == dst_port && (dg & (ND_SINGLE | ND_RETIRE | ND_ASLEEP)) && nd_sock(dg) != ND_NOSOCK) s = (int)nd_sock(dg);
            } else {
                rc = classify(node, s);
            }
            if (rc == RC_RETIRED) { found = true; retire = true; }
            else if (rc == RC_VIRTUAL) {     /* its sleeping owner's loop head will pop and reject it: counted, not queued */
                found = true; queued_virtually = true;
                sockv_add(R, (u32)s, K.size);
            }
            else if (rc == RC_TABLES) {
                if (s < 0) s = sock_find(R, node, K.dst_ip, dst_port);
                found = s >= 0;
                if (fast_ok && found) {
                    const Sock& S = R.sock[s];
                    const u16 w = S.wait_head;
                    if (w != NIL16 && S.q_len == 0 && R.pcb[w].wnext == NIL16) {
                        const Pcb& P = R.pcb[w];
                        const u64 ins = R.m.code[P.pc];
                        const u32 wop = (u32)(ins & 0xFF);
                        if ((wop == OP_RECV_FILTER || wop == OP_RECV_FILTER_LOG) && P.phase == 1 &&
                            R.node[node].epoch == P.r[(u32)(ins >> 8) & 3]) {
                            const itsb_filter F = R.m.filters[(u32)(ins >> 16) & 0xFF];
                            retire = !filter_passes(R, F, K.tag, K.f0, R.m.nodes[node].host, P);
                        }
                    }
                }
            }
        }
        const u32 n_found = (u32)popc(w_ballot(found));
        const u32 n_retired = (u32)popc(w_ballot(retire));
        const bool acc = found && !retire && !queued_virtually;      /* recipients that really enqueue */
        /* a copy goes onto each accepting socket: into its inline head slot when its queue is empty, else an arena node */
        const bool need_node = acc && (!inl || R.sock[s].q_len != 0);
        const u32 amask = w_ballot(need_node);
        const u32 n_acc = (u32)popc(amask);
        const u32 n_valid = (u32)popc(w_ballot(valid));
        if (R.trace)
            trace_deliveries(H, i - base, node, valid, (K.size & 0x7FFFFFFFu) | (found ? 0x80000000u : 0u), n_valid);
        u32 atop = H->arena_free_top;
        if (atop < n_acc) { fail(H, ITSB_ERR_CAP_ARENA, k); return false; }
        u32 nwait = 0;
        if (acc) {
            ArenaNode A;
            A.next = NIL32; A.src_ip = K.src_ip; A.ports = K.ports; A.size = K.size;
            A.tag = K.tag; A.f0 = K.f0; A.f1 = K.f1; A.pad = K.dst_ip;
            Sock& S = R.sock[s];
            if (need_node) {
                u32 rank = (u32)popc(amask & lanemask_lt());
                u32 an = R.arena_free[atop - 1 - rank];
                R.arena[an] = A;
                if (S.q_tail == NIL32) S.q_head = an; else R.arena[S.q_tail].next = an;
                S.q_tail = an;
            } else {
                R.sockpkt[s] = A;
            }
            S.q_len += 1;
            for (u16 w = S.wait_head; w != NIL16; w = R.pcb[w].wnext) nwait += 1;
            if (nwait == 0) node_digest_enqueued(R, node, S.port);
        }
        if (n_acc) H->arena_free_top = atop - n_acc;
        /* _enqueue -> packet signal turn_on: resume this socket's waiters (normally the one blocked in recv) */
        u32 total, total_seq;
        u32 before = w_excl_scan(nwait, &total);                               /* ring position: real wake-ups */
        u32 before_seq = w_excl_scan(nwait + (retire ? 1u : 0u), &total_seq);  /* counter value: every waiter  */
        const bool ring_ok = total == 0 || (H->nq_spill_count == 0 && (NQ_CAP - (H->nq_tail - H->nq_head)) >= total);
        if (ring_ok) {
            /* every lane writes its socket's resumes straight into the ring slots its prefix rank assigns */
            if (nwait) {
                Sock& S = R.sock[s];
                u32 j = 0;
                u16 w = S.wait_head;
                while (w != NIL16) {
                    Pcb& P = R.pcb[w];
                    u16 nx = P.wnext;
                    P.wnext = NIL16; P.wprev = NIL16; P.wsig = WS_NONE; P.state = PS_RESUMING;
                    u32 q = (H->nq_tail + before + j) & (NQ_CAP - 1);
                    R.nq[2 * q] = H->seq + before_seq + j;
                    R.nq[2 * q + 1] = ev_pack(ITSB_EV_WAKE, w, P.gen, 0);
                    j += 1;
                    w = nx;
                }
                S.wait_head = NIL16; S.wait_tail = NIL16;
                node_digest_woken(R, node, S.port);
            }
            w_sync();
            /* a model that keeps connection records: the retired recipients' recv() would have correlated the packet
               (overrides.py:135-147); done here, one socket after the other in recipient order, each under the counter
               value its wake-up would have had (the order key of the records it logs) */
            if (n_retired && H->conntab) conn_retired(H, w_ballot(retire), (u32)s, H->seq + before_seq, k);
            if (total) H->nq_tail += total;
            if (total_seq) H->seq += total_seq;
        } else {
            /* ring nearly full: the same pushes, in recipient order, through the spilling path (warp-uniform), each
               recipient's counter values -- a retired wake-up's or its waiters' -- taken in that order */
#if defined(__CUDACC__)
#pragma unroll 1
#endif
            for (int l = 0; l < ITSB_W; ++l) {
                u32 nl = w_shfl(nwait, l);
                u32 sl = w_shfl((u32)s, l);
                const bool rl = w_shfl(retire ? 1u : 0u, l) != 0;
                if (rl) {
                    if (H->conntab) conn_retired(H, 1u << l, (u32)s, H->seq, k);
                    H->seq += 1;
                }
                if (nl != 0) sock_resume_waiters(H, sl);
            }
        }
#if ITSB_W == 1
        q_valid += n_valid; q_found += n_found; q_retired += n_retired;     /* flushed after the pass */
#else
        R.stats[ITSB_TM_KIND0 + ITSB_EV_DELIVER] += n_valid;
        R.stats[ITSB_TM_DELIVERED] += n_found;
        R.stats[ITSB_TM_DROPPED] += n_valid - n_found;
        if (n_retired) {
            R.stats[ITSB_TM_KIND0 + ITSB_EV_WAKE] += n_retired;
            R.stats[ITSB_TM_PACKETS_RECEIVED] += n_retired;
            R.stats[ITSB_TM_BYTES_RECEIVED] += (u64)n_retired * K.size;
        }
#endif
        w_sync();
        return true;
    };

#if ITSB_W == 1
    if (R.trace == nullptr) {
        /* A lane per replica, two phases per 64 recipients.  (1) every recipient is classified from its digest alone:
           a tight loop over 4-byte words whose sectors are fetched together, in which the lanes of a warp that are
           delivering stay in step.  (2) the recipients left for the tables -- typically the one socket that accepts
           and the daemons of sleeping workstations -- are taken in recipient order; the counter values the retired
           ones in between consume are accounted for as the pass reaches them, so every wake-up gets the value it
           would have got with the recipients taken strictly one after the other. */
#if defined(__CUDACC__)
        if (count > 8) {
            for (u32 o = 0; o < count; o += 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(R.ndig + first + o));
            u32 listed_class = 0;       /* does some filter look at the remembered packets for this tag? then the Bloom words too */
            for (u32 f = 0; f < R.m.n_filters && f <= ND_MAX_FILTER; ++f) listed_class |= R.m.filters[f].tags_if_listed & tag_bit;
            if (listed_class)
                for (u32 o = 0; o < count; o += 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(R.nmask + first + o));
        }
#endif
        for (u32 c0 = 0; c0 < count; c0 += 64) {
            const u32 cn = count - c0 < 64 ? count - c0 : 64;
            u64 slow = 0, retired = 0;
            u32 dropped = 0, queued = 0;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
            for (u32 j = 0; j < cn; ++j) {
                int sj = -1;
                const u32 rc = classify(first + c0 + j, sj);
                if (rc == RC_TABLES) slow |= 1ull << j;
                else if (rc == RC_RETIRED) retired |= 1ull << j;
                else if (rc == RC_VIRTUAL) { sockv_add(R, (u32)sj, K.size); queued += 1; }
                else dropped += 1;
            }
            if (H->status != 0) return;
            const u32 n_ret = (u32)popc64(retired);
            q_valid += dropped + n_ret + queued; q_found += n_ret + queued; q_retired += n_ret;
            while (slow != 0) {
                const u32 j = (u32)ctz64(slow);
                slow &= slow - 1;
                const u64 below = (1ull << j) - 1;
                const u32 r = (u32)popc64(retired & below);     /* retired recipients before this one */
                if (r) { H->seq += r; retired &= ~below; }
                if (!pass(c0 + j, true)) return;
            }
            const u32 r = (u32)popc64(retired);
            if (r) H->seq += r;
        }
    } else {
        for (u32 base = 0; base < count; base += ITSB_W) if (!pass(base, false)) return;
    }
    if (q_valid) {
        stat_add(R, ITSB_TM_KIND0 + ITSB_EV_DELIVER, q_valid);
        if (q_found) stat_add(R, ITSB_TM_DELIVERED, q_found);
        if (q_valid != q_found) stat_add(R, ITSB_TM_DROPPED, q_valid - q_found);
        if (q_retired) {
            stat_add(R, ITSB_TM_KIND0 + ITSB_EV_WAKE, q_retired);
            stat_add(R, ITSB_TM_PACKETS_RECEIVED, q_retired);
            stat_add(R, ITSB_TM_BYTES_RECEIVED, (u64)q_retired * K.size);
        }
    }
#else
    for (u32 base = 0; base < count; base += ITSB_W) if (!pass(base, false)) return;
#endif
    arena_release(R, k);
}

/* ---- interpreter ------------------------------------------------------------------------------------------------------------- */
ITSB_FN u32 vol_index(u32 s) {
    u32 i = s - 0x10u;
    return i < VOL_COUNT ? i : VOL_COUNT - 1;
}

/* operand value: r0..r3, or a scratch register (packet fields, list entry, node constants, ...) */
ITSB_FN u32 src_val(Rep& R, const Pcb& P, u32 s) { return s < 4 ? P.r[s] : R.hdr->vol[vol_index(s)]; }

ITSB_FN void set_pkt_regs(Rep& R, const ArenaNode& A) {
    u32* v = R.hdr->vol;
    v[SRC_PKT_SRC_IP - 0x10] = A.src_ip; v[SRC_PKT_SRC_PORT - 0x10] = A.ports & 0xFFFF;
    v[SRC_PKT_SIZE - 0x10] = A.size; v[SRC_PKT_TAG - 0x10] = A.tag;
    v[SRC_PKT_F0 - 0x10] = A.f0; v[SRC_PKT_F1 - 0x10] = A.f1; v[SRC_PKT_DST_IP - 0x10] = A.pad;
}

enum { XR_CONTINUE = 0, XR_STOP = 1 };

}  // namespace itsb
#include "itsb_machine.cuh"

#if defined(ITSB_HIDDEN_FAST)
namespace itsb {
/* One activation (start or resume) of select()'s per-signal helper, without the interpreter: the program is
 * `signal.wait(); common.turn_on()` (ir.py add_hidden_programs: SIG_WAIT r0 / SIG_ON r1 / EXIT), so an activation either
 * parks on the signal again or turns the common signal on and ends.  Same calls, in the same order, as the interpreted
 * program makes: every counter value and queue position is unchanged.  Compiled into the in-place kernel (five of the
 * seven heap pops behind a received packet of a shipped-itsim model are such helper hops) and the host emulation. */
ITSB_COLD void hidden_wait_one(Hdr* H, u32 p) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    P.state = PS_RUNNING;
    if (H->status != 0) return;
    if (sig_wait(R, p, P.r[0] & 0xFFFF, 0xFFFF)) return;
    sig_on(R, P.r[1] & 0xFFFF);
    proc_exit(H, p);
}
}  // namespace itsb
#endif
namespace itsb {

ITSB_COLD int exec_logged(Hdr* H, u32 p, u64 ins);   /* instructions of models that keep connection records */

/* Every instruction that is not on the per-packet path of a daemon: out of line. */
ITSB_COLD int exec_cold(Hdr* H, u32 p, u64 ins) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    const u32 op = (u32)(ins & 0xFF), a = (u32)((ins >> 8) & 0xFF), b = (u32)((ins >> 16) & 0xFF),
              c = (u32)((ins >> 24) & 0xFF), imm = (u32)(ins >> 32);
    switch (op) {
    case OP_NOP:
        P.pc += 1;
        return XR_CONTINUE;
    case OP_EXIT:
        proc_exit(H, p);
        return XR_STOP;
    case OP_LI:
        P.r[a & 3] = imm; P.pc += 1;
        return XR_CONTINUE;
    case OP_MOV:
        P.r[a & 3] = src_val(R, P, b); P.pc += 1;
        return XR_CONTINUE;
    case OP_ADDI:
        P.r[a & 3] = src_val(R, P, b) + imm; P.pc += 1;
        return XR_CONTINUE;
    case OP_DRAWI:
        P.r[a & 3] = (u32)(int32_t)(long long)draw_dist(H, &R.m.dists[imm]);
        P.pc += 1;
        return XR_CONTINUE;
    case OP_DRAW_ADDI: {
        double v = (double)(int32_t)src_val(R, P, b) + draw_dist(H, &R.m.dists[imm]);
        P.r[a & 3] = (u32)(int32_t)(long long)v;
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_ADVANCE:
    case OP_ADVANCE_C: {
        double d = (op == OP_ADVANCE) ? draw_dist(H, &R.m.dists[imm & 0xFFFF]) : R.m.consts[imm & 0xFFFF];
        if (d < 0.0) { fail(H, ITSB_EXC_NEGATIVE_DELAY, P.pc); return XR_STOP; }
        P.timer_tok = (u16)(P.timer_tok + 1);
        int slot = ev_push_inl(R, H->now + d, ev_pack(ITSB_EV_TIMER, p, P.gen, P.timer_tok));
        P.timer_ev = slot < 0 ? NIL16 : (u16)slot;
        P.state = PS_TIMER;
        return XR_STOP;  /* pc stays on the ADVANCE; the TIMER event moves it on */
    }
    case OP_OPEN:
        if (sock_open(H, p, imm & 0xFFFF) < 0) return XR_STOP;
        P.pc += 1;
        return XR_CONTINUE;
    case OP_CLOSE:
        sock_close(H, p);
        P.pc += 1;
        return XR_CONTINUE;
    case OP_SEND: {
        const u32* v = H->vol;
        u32 dst_ip, dst_port = imm & 0xFFFF;
        switch (a) {
        case DST_BCAST: dst_ip = R.m.nets[R.m.nodes[P.node].net].bcast; break;
        case DST_PKT_SRC: dst_ip = v[SRC_PKT_SRC_IP - 0x10]; dst_port = v[SRC_PKT_SRC_PORT - 0x10]; break;
        case DST_ENT: dst_ip = v[SRC_ENT_IP - 0x10]; dst_port = v[SRC_ENT_PORT - 0x10]; break;
        case DST_SELF: dst_ip = R.m.nodes[P.node].addr; break;
        default: dst_ip = P.r[2]; dst_port = P.r[3] & 0xFFFF; break;
        }
        u32 f0s = (imm >> 16) & 0xFF, f1s = (imm >> 24) & 0xFF;
        if (P.sock == NIL16) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return XR_STOP; }
        net_send(H, P.node, R.sock[P.sock].port, dst_ip, dst_port, src_val(R, P, b), c,
                 f0s == SRC_NONE ? 0 : src_val(R, P, f0s), f1s == SRC_NONE ? 0 : src_val(R, P, f1s));
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_FORWARD: {
        /* world_socket.send(dest, msg.byte_size, msg.payload) from another attachment of this machine
           (malware_simple.py:101-107): source = (address of node imm[0:16], port imm[16:32]), dest = (r2, r3) */
        const u32* v = H->vol;
        net_send(H, imm & 0xFFFF, imm >> 16, P.r[2], P.r[3] & 0xFFFF, v[SRC_PKT_SIZE - 0x10], v[SRC_PKT_TAG - 0x10],
                 v[SRC_PKT_F0 - 0x10], v[SRC_PKT_F1 - 0x10]);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_NODE_SLEEP: {
        NodeDyn& N = R.node[P.node];
        N.awake = 0; N.epoch += 1;
        node_digest_epoch(R, P.node);
        H->vol[SRC_NODE_EPOCH - 0x10] = N.epoch;
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_NODE_WAKE:
        node_wake(H, P.node);
        H->vol[SRC_NODE_EPOCH - 0x10] = R.node[P.node].epoch;
        P.pc += 1;
        return XR_CONTINUE;
    case OP_SPAWN: {
        int child = proc_spawn(H, imm, P.node, P.r, 0.0);
        if (child < 0) return XR_STOP;
        if (a != SRC_NONE) P.r[a & 3] = handle_of(R, (u32)child);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_THROW: {
        u32 h = src_val(R, P, a);
        ev_push_inl(R, H->now, ev_pack(ITSB_EV_THROW, h & PROC_MASK, (h >> 16) & GEN_MASK, imm));
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_LIST_CLEAR:
        list_free_all(R, P);
        R.pmask[p] = 0;
        P.pc += 1;
        return XR_CONTINUE;
    case OP_LIST_PUSH: {
        u32 n = arena_alloc(R);
        if (n == NIL32) return XR_STOP;
        const u32* v = H->vol;
        ArenaNode A;
        A.next = NIL32; A.src_ip = v[SRC_PKT_SRC_IP - 0x10]; A.ports = v[SRC_PKT_SRC_PORT - 0x10];
        A.size = v[SRC_PKT_SIZE - 0x10]; A.tag = v[SRC_PKT_TAG - 0x10];
        A.f0 = v[SRC_PKT_F0 - 0x10]; A.f1 = v[SRC_PKT_F1 - 0x10]; A.pad = 0;
        R.arena[n] = A;
        if (P.list_tail != NIL32) R.arena[P.list_tail].next = n;
        if (P.list_tail == NIL32) P.list_head = n;
        P.list_tail = n;
        R.pmask[p] |= 1ull << (A.f0 & 63);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_LIST_TAKE: {
        u32 key = src_val(R, P, a);
        u32 prev = NIL32, n = P.list_head;
        bool found = false;
        while (n != NIL32) {
            const ArenaNode A = R.arena[n];
            if (A.f0 == key) {
                H->vol[SRC_ENT_IP - 0x10] = A.src_ip; H->vol[SRC_ENT_PORT - 0x10] = A.ports & 0xFFFF;
                if (prev == NIL32) P.list_head = A.next;
                else R.arena[prev].next = A.next;
                if (P.list_tail == n) P.list_tail = prev;
                arena_release(R, n);
                found = true;
                break;
            }
            prev = n; n = A.next;
        }
        if (found) {     /* the Bloom filter of what is left (two remembered packets may share a bit) */
            u64 mk = 0;
            for (u32 q = P.list_head; q != NIL32; q = R.arena[q].next) mk |= 1ull << (R.arena[q].f0 & 63);
            R.pmask[p] = mk;
        }
        P.pc = found ? (u16)(P.pc + 1) : (u16)imm;
        return XR_CONTINUE;
    }
    case OP_STAT:
        stat_add(R, ITSB_TM_USER(imm), 1);
        P.pc += 1;
        return XR_CONTINUE;
    default:
        return op >= OP_RECV_LOG ? exec_logged(H, p, ins) : exec_machine(H, p, ins);
    }
}

/* thread / process handles of an activation (shipped-itsim models) */
ITSB_COLD void thread_regs(Hdr* H, u32 p) {
    Rep R; rebind(R, H);
    const u32 th = R.pcbx[p].thread;
    if (th == NIL16) return;
    const u32 tk = R.thr[th].task;
    u32* v = H->vol;
    v[SRC_SELF_THREAD - 0x10] = th | ((u32)R.thr[th].gen << 16);
    v[SRC_SELF_TASK - 0x10] = tk | ((u32)R.task[tk].gen << 16);
    v[SRC_PID - 0x10] = R.task[tk].pid;
}

/* Runs process p from its pc until it blocks or ends.  Inline: the receive loop heads, waits and branches a daemon
 * or a query executes per packet; everything else goes through exec_cold. */
ITSB_FN void run_process(Rep& R, u32 p) {
    Hdr* H = R.hdr;
    Pcb& P = R.pcb[p];
    P.state = PS_RUNNING;
    {   /* node constants of this activation */
        const itsb_node N = R.m.nodes[P.node];
        u32* v = H->vol;
        v[SRC_NODE_ADDR - 0x10] = N.addr; v[SRC_NODE_INDEX - 0x10] = P.node; v[SRC_NODE_HOST - 0x10] = N.host;
        v[SRC_NODE_EPOCH - 0x10] = R.node[P.node].epoch; v[SRC_NODE_BCAST - 0x10] = R.m.nets[N.net].bcast;
        v[SRC_SELF - 0x10] = handle_of(R, p); v[SRC_EXC - 0x10] = P.exc; v[SRC_ZERO - 0x10] = 0;
        v[SRC_TAGS - 0x10] = P.tags;
        if (R.max_sig != 0) thread_regs(H, p);
    }
    /* the instructions handled inline below (all numbered under 32); everything else is one test away from exec_cold */
    const u32 INLINE_OPS = (1u << OP_RECV_FILTER) | (1u << OP_RECV) | (1u << OP_WAIT_AWAKE) | (1u << OP_JMP) |
                           (1u << OP_JEQI) | (1u << OP_JNEI) | (1u << OP_JEQ) | (1u << OP_JNE) | (1u << OP_JASLEEP);
    if (H->status != 0) return;
    for (;;) {
        const u64 ins = R.m.code[P.pc];
        const u32 op = (u32)(ins & 0xFF), a = (u32)((ins >> 8) & 0xFF), b = (u32)((ins >> 16) & 0xFF),
                  imm = (u32)(ins >> 32);
        if (op >= 32 || ((INLINE_OPS >> op) & 1) == 0) {
            if (exec_cold(H, p, ins) != XR_CONTINUE || H->status != 0) return;
        } else if (op == OP_RECV_FILTER) {
            if (P.sock == NIL16) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return; }
            NodeDyn& N = R.node[P.node];
            Sock& S = R.sock[P.sock];
            const itsb_filter F = R.m.filters[b];
            bool handed = false;
            /* the node's wake state cannot change inside this loop (only NODE_SLEEP / NODE_WAKE of the blinking process
               move it): read once, together with the socket, instead of after the packet has arrived from HBM */
            const u32 n_epoch = N.epoch, n_awake = N.awake;
            for (;;) {
                if (P.phase == 0) {               /* with ws.awake(): wait_until_awake, remember the wake epoch */
                    if (!n_awake) {
                        waiters_append(R, N.wake_head, N.wake_tail, p, (u16)(WS_NODE | P.node));
                        P.state = PS_WAIT;
                        node_digest_asleep(R, P.node, p);
                        return;
                    }
                    const SockV V = sockv_take(R, P.sock);
                    if (V.count != 0) {           /* what was queued virtually while it slept: popped and rejected here */
                        stat_add(R, ITSB_TM_PACKETS_RECEIVED, V.count);
                        stat_add(R, ITSB_TM_BYTES_RECEIVED, V.bytes);
                    }
                    P.r[a & 3] = n_epoch;
                    P.phase = 1;
                }
                if (S.q_len == 0) {               /* socket.recv() */
                    waiters_append(R, S.wait_head, S.wait_tail, p, (u16)(WS_SOCK | P.sock));
                    P.state = PS_WAIT;
                    node_digest_parked(R, S.node, P.sock, p, b, a);
                    return;
                }
                const ArenaNode A = sock_pop(R, S);
                P.phase = 0;
                if (n_epoch != P.r[a & 3]) { P.pc = (u16)(imm & 0xFFFF); break; }      /* FellAsleep */
                if (filter_passes(R, F, A.tag, A.f0, R.m.nodes[P.node].host, P)) {
                    set_pkt_regs(R, A);
                    handed = true;
                    break;
                }
            }
            if (handed) P.pc += 1;
        } else if (op == OP_RECV) {
            /* while queue.empty(): signal.wait()  (overrides.py:135-138) */
            if (P.sock == NIL16) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return; }
            Sock& S = R.sock[P.sock];
            if (S.q_len == 0) {
                waiters_append(R, S.wait_head, S.wait_tail, p, (u16)(WS_SOCK | P.sock));
                P.state = PS_WAIT;
                node_digest_waits(R, S.node, P.sock);
                return;
            }
            const ArenaNode A = sock_pop(R, S);
            set_pkt_regs(R, A);
            P.pc += 1;
        } else if (op == OP_WAIT_AWAKE) {
            NodeDyn& N = R.node[P.node];
            if (!N.awake) {
                waiters_append(R, N.wake_head, N.wake_tail, p, (u16)(WS_NODE | P.node));
                P.state = PS_WAIT;
                return;
            }
            if (a != SRC_NONE) P.r[a & 3] = N.epoch;
            P.pc += 1;
        } else if (op == OP_JMP) {
            P.pc = (u16)imm;

        } else if (op == OP_JEQI || op == OP_JNEI) {
            const bool eq = src_val(R, P, a) == (imm >> 16);
            P.pc = (eq == (op == OP_JEQI)) ? (u16)(imm & 0xFFFF) : (u16)(P.pc + 1);
        } else if (op == OP_JEQ || op == OP_JNE) {
            const bool eq = src_val(R, P, a) == src_val(R, P, b);
            P.pc = (eq == (op == OP_JEQ)) ? (u16)imm : (u16)(P.pc + 1);
        } else {   /* OP_JASLEEP */
            P.pc = (R.node[P.node].epoch != src_val(R, P, a)) ? (u16)imm : (u16)(P.pc + 1);
        }
    }
}

/* An exception arrives at a blocked process (greenlet.throw): leave whatever it was parked in, then run its handler.
 * greensim would leave a "ghost" wake-up / queue entry for non-Interrupt exceptions; ghosts are no-ops, so the
 * engine removes them instead (same model-visible order).  Returns the process to run, or NIL32. */
ITSB_COLD u32 deliver_exception(Hdr* H, u32 p, u32 exc) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    if (P.state == PS_TIMER) {
        /* advance() interrupted: its wake-up is cancelled (pool) or left to expire as a dead hop (queue / overflow) */
        if (P.timer_ev != NIL16) ev_release(R, P.timer_ev);
        P.timer_ev = NIL16;
        P.timer_tok = (u16)(P.timer_tok + 1);
    } else if (P.state == PS_WAIT && (P.wsig & 0xC000) != WS_SIG) {
        if (P.wsig & WS_NODE) {
            const u32 nd = P.wsig & 0x3FFF;
            NodeDyn& N = R.node[nd];
            waiters_unlink(R, N.wake_head, N.wake_tail, p);
            if (R.ndig[nd] & ND_ASLEEP) node_digest_refresh(H, nd);     /* it may be the sleeper the digest speaks of */
        }
        else if (P.wsig & WS_SOCK) {
            Sock& S = R.sock[P.wsig & 0x3FFF];
            waiters_unlink(R, S.wait_head, S.wait_tail, p);
            node_digest_refresh(H, S.node);
        }
    }
    exc = machine_unwind(R, p, exc);
    P.exc = (u16)exc;
    P.phase = 0;
    const u64 ins = R.m.code[P.pc];
    u32 handler = (u32)(ins >> 48);
    if (handler == 0) {
        if (!(exc & EXC_INTERRUPT_CLASS)) { fail(H, ITSB_EXC_UNCAUGHT, ((u64)exc << 32) | P.pc); return NIL32; }
        handler = 0;    /* code[0..1] = CLOSE; EXIT -- an uncaught Interrupt ends the process silently */
    }
    P.pc = (u16)handler;
    return p;
}

/* Set-up time draws: `infected = next(distribution(workstation_list))` (malware_simple.py:114) happens before the
 * run, from the replica's own stream; the installed process i sits in slot i of a fresh replica. */
ITSB_COLD void replica_setup(Hdr* H) {
    Rep R; rebind(R, H);
    H->setup_done = 1;
    u32 drawn = 0;
    for (u32 i = 0; i < R.m.n_init; ++i) {
        const itsb_initproc I = R.m.init[i];
        if (I.mode == ITSB_INIT_ADD_DRAWN) {
            drawn = I.node + (u32)(long long)draw_dist(H, &R.m.dists[I.delay_const]);
            R.pcb[i].node = (u16)drawn;
        } else if (I.mode == ITSB_INIT_ADD_SAME) {
            R.pcb[i].node = (u16)drawn;
        }
    }
}

ITSB_FN u64 events_counted(Rep& R) {
    u64 n = 0;
    for (u32 k = 0; k < ITSB_NUM_KINDS; ++k) n += stat_get(R, ITSB_TM_KIND0 + k);
    return n;
}

/* Simulator.run(duration): pops (t, seq)-ordered events until the stop event.  Returns the number
 * of simulated events this call processed. */
ITSB_FN u64 run_replica(Rep& R, double duration) {
    Hdr* H = R.hdr;
    if (H->status != 0) return 0;
    if (!H->setup_done) replica_setup(H);
    const u64 events_before = events_counted(R);
    /* run(duration): a finite duration schedules the stop event first (it takes a counter value); run() without
       one ends when the heap is empty and leaves the clock where it is */
    const double t_stop = H->now + duration;
    const u64 t_stop_bits = dbits(t_stop);
    const bool finite = t_stop_bits < T_EMPTY;
    const u32 stop_seq = finite ? H->seq : NIL32;
    if (finite) H->seq += 1;
    for (;;) {
        if (H->status != 0) break;
        u64 tb; u32 seq;
        const int src = ev_peek(R, tb, seq);
        if (src == SRC_NONE_LEFT && !finite) break;
        if (src == SRC_NONE_LEFT || tb > t_stop_bits || (tb == t_stop_bits && seq > stop_seq)) {
            H->now = t_stop;
            count_event(R, ITSB_EV_STOP, 0, 0, 0);
            break;
        }
        const u32 data = ev_take(R, src);
        H->now = bitsd(tb);
        H->cur_seq = seq;
        const u32 kind = ev_kind(data), idx = ev_idx(data);
        u32 runnable = NIL32, deliver = NIL32;
        if (kind == ITSB_EV_DELIVER) {
            deliver = ev_pkt(data);
        } else if (kind == ITSB_EV_WAKE || kind == ITSB_EV_TIMER || kind == ITSB_EV_PROC_START) {
            Pcb& P = R.pcb[idx];
            bool live = P.state != PS_FREE && P.gen == ev_gen(data);   /* else a ghost hop: target finished */
            if (live) {
                if (kind == ITSB_EV_TIMER) {
                    live = P.state == PS_TIMER && (u32)(P.timer_tok & 0xFF) == ev_exc(data);
                    if (live) { P.timer_ev = NIL16; P.pc += 1; }
                } else if (kind == ITSB_EV_WAKE) {
                    live = P.state == PS_RESUMING;
                } else {
                    live = P.state == PS_UNSTARTED;
                }
            }
            if (live) {
#if defined(ITSB_HIDDEN_FAST)
                if (P.prog == PROG_HIDDEN_WAIT_ONE) hidden_wait_one(H, idx); else
#endif
                {
                if (!is_hidden_prog(P.prog)) count_event(R, kind, P.prog, P.node, 0);
                runnable = idx;
                }
            }
        } else if (kind == ITSB_EV_TX_BEGIN) {
            on_tx_begin(H, ev_pkt(data));
        } else if (kind == ITSB_EV_TX_END) {
            if (on_tx_end(H, ev_pkt(data))) deliver = ev_pkt(data);
        } else if (kind == ITSB_EV_THROW) {
            Pcb& P = R.pcb[idx];
            if (P.state != PS_FREE && P.gen == ev_gen(data)) {
                if (!is_hidden_prog(P.prog) && !(ev_exc(data) & EXC_HIDDEN))
                    count_event(R, ITSB_EV_THROW, P.prog, P.node, ev_exc(data));
                if (P.state == PS_UNSTARTED) proc_exit(H, idx);   /* never ran: dies on the spot */
                else runnable = deliver_exception(H, idx, ev_exc(data));
#if defined(ITSB_HIDDEN_FAST)
                /* a helper whose handler is the program's EXIT (CleanUp / CancelBalk: `except ...: pass`) ends here */
                if (runnable != NIL32 && is_hidden_prog(P.prog) && (u32)(R.m.code[P.pc] & 0xFF) == OP_EXIT) {
                    proc_exit(H, idx);
                    runnable = NIL32;
                }
#endif
            }
        } else {
            fail(H, ITSB_ERR_BAD_OPCODE, data);
        }
        if (deliver != NIL32) on_deliver(R, deliver);      /* the delivery pass's only call site */
        if (runnable != NIL32) run_process(R, runnable);   /* the interpreter's only call site */
    }
    return events_counted(R) - events_before;
}

/* ---- the same loop, cut in two for the phase-aligned kernel (itsb_phased.cu) ------------------------------------------
 * run_replica == run_begin; while (run_pop(runnable)) if (runnable) run_process(runnable); run_events.  The phased kernel
 * lets every warp of the CTA do its pop (+ network work) first and its process activation second, with a CTA barrier
 * in between, so that warps of one SM execute the same half of the code at the same time. */
struct RunCtx { double t_stop; u64 t_stop_bits; u64 events_before; u32 stop_seq; bool finite; bool active; };

ITSB_FN void run_begin(Rep& R, double duration, RunCtx& C) {
    Hdr* H = R.hdr;
    C.events_before = 0;
    C.active = H->status == 0;
    if (!C.active) return;
    if (!H->setup_done) replica_setup(H);
    C.events_before = events_counted(R);
    C.t_stop = H->now + duration;
    C.t_stop_bits = dbits(C.t_stop);
    C.finite = C.t_stop_bits < T_EMPTY;
    C.stop_seq = C.finite ? H->seq : NIL32;
    if (C.finite) H->seq += 1;
}

/* One pop.  Returns false when this run() is over.  Network events (TX_BEGIN, TX_END, the delivery pass) are handled
 * here; a process to activate is handed back in `runnable` (NIL32: none). */
ITSB_FN bool run_pop(Rep& R, const RunCtx& C, u32& runnable) {
    Hdr* H = R.hdr;
    runnable = NIL32;
    if (!C.active || H->status != 0) return false;
    u64 tb; u32 seq;
    const int src = ev_peek(R, tb, seq);
    if (src == SRC_NONE_LEFT && !C.finite) return false;
    if (src == SRC_NONE_LEFT || tb > C.t_stop_bits || (tb == C.t_stop_bits && seq > C.stop_seq)) {
        H->now = C.t_stop;
        count_event(R, ITSB_EV_STOP, 0, 0, 0);
        return false;
    }
    const u32 data = ev_take(R, src);
    H->now = bitsd(tb);
    H->cur_seq = seq;
    const u32 kind = ev_kind(data), idx = ev_idx(data);
    u32 deliver = NIL32;
    if (kind == ITSB_EV_DELIVER) {
        deliver = ev_pkt(data);
    } else if (kind == ITSB_EV_WAKE || kind == ITSB_EV_TIMER || kind == ITSB_EV_PROC_START) {
        Pcb& P = R.pcb[idx];
        bool live = P.state != PS_FREE && P.gen == ev_gen(data);
        if (live) {
            if (kind == ITSB_EV_TIMER) {
                live = P.state == PS_TIMER && (u32)(P.timer_tok & 0xFF) == ev_exc(data);
                if (live) { P.timer_ev = NIL16; P.pc += 1; }
            } else if (kind == ITSB_EV_WAKE) {
                live = P.state == PS_RESUMING;
            } else {
                live = P.state == PS_UNSTARTED;
            }
        }
        if (live) {
            if (!is_hidden_prog(P.prog)) count_event(R, kind, P.prog, P.node, 0);
            runnable = idx;
        }
    } else if (kind == ITSB_EV_TX_BEGIN) {
        on_tx_begin(H, ev_pkt(data));
    } else if (kind == ITSB_EV_TX_END) {
        if (on_tx_end(H, ev_pkt(data))) deliver = ev_pkt(data);
    } else if (kind == ITSB_EV_THROW) {
        Pcb& P = R.pcb[idx];
        if (P.state != PS_FREE && P.gen == ev_gen(data)) {
            if (!is_hidden_prog(P.prog) && !(ev_exc(data) & EXC_HIDDEN))
                count_event(R, ITSB_EV_THROW, P.prog, P.node, ev_exc(data));
            if (P.state == PS_UNSTARTED) proc_exit(H, idx);
            else runnable = deliver_exception(H, idx, ev_exc(data));
        }
    } else {
        fail(H, ITSB_ERR_BAD_OPCODE, data);
    }
    if (deliver != NIL32) on_deliver(R, deliver);
    return true;
}

ITSB_FN u64 run_events(Rep& R, const RunCtx& C) { return C.active ? events_counted(R) - C.events_before : 0; }

/* ---- the loop cut at the pop, for the kernel that sorts replicas by what they do next (itsb_sorted.cu) ---------------
 * run_replica == run_begin; while (run_take(ev)) run_handle(ev).  The event is taken from the queue first and handled in
 * a later step; nothing touches the replica in between, so the order of everything inside it is the original one. */
struct PendingEvent { u64 tb; u32 seq; u32 data; };

/* Takes the next event (or ends the run: stop event / empty heap).  Returns false when this run() is over. */
ITSB_FN bool run_take(Rep& R, const RunCtx& C, PendingEvent& E) {
    Hdr* H = R.hdr;
    if (!C.active || H->status != 0) return false;
    u64 tb; u32 seq;
    const int src = ev_peek(R, tb, seq);
    if (src == SRC_NONE_LEFT && !C.finite) return false;
    if (src == SRC_NONE_LEFT || tb > C.t_stop_bits || (tb == C.t_stop_bits && seq > C.stop_seq)) {
        H->now = C.t_stop;
        count_event(R, ITSB_EV_STOP, 0, 0, 0);
        return false;
    }
    E.data = ev_take(R, src);
    E.tb = tb;
    E.seq = seq;
#if defined(__CUDACC__) && defined(ITSB_LANES)
    /* the event is handled in a later round (k_run_vote) or by another warp (k_run_pool): what its handler reads first
       -- the process control block, or the packet -- starts its way from HBM now */
    {
        const u32 kind = ev_kind(E.data);
        const void* first = (kind == ITSB_EV_TX_BEGIN || kind == ITSB_EV_TX_END || kind == ITSB_EV_DELIVER)
                                ? (const void*)(R.arena + ev_pkt(E.data)) : (const void*)(R.pcb + ev_idx(E.data));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(first));
    }
#endif
    return true;
}

/* What the pending event does (the body of run_pop after the pop), including the activation it leads to. */
ITSB_FN void run_handle(Rep& R, const PendingEvent& E) {
    Hdr* H = R.hdr;
    const u32 data = E.data;
    H->now = bitsd(E.tb);
    H->cur_seq = E.seq;
    const u32 kind = ev_kind(data), idx = ev_idx(data);
    u32 runnable = NIL32, deliver = NIL32;
    if (kind == ITSB_EV_DELIVER) {
        deliver = ev_pkt(data);
    } else if (kind == ITSB_EV_WAKE || kind == ITSB_EV_TIMER || kind == ITSB_EV_PROC_START) {
        Pcb& P = R.pcb[idx];
        bool live = P.state != PS_FREE && P.gen == ev_gen(data);
        if (live) {
            if (kind == ITSB_EV_TIMER) {
                live = P.state == PS_TIMER && (u32)(P.timer_tok & 0xFF) == ev_exc(data);
                if (live) { P.timer_ev = NIL16; P.pc += 1; }
            } else if (kind == ITSB_EV_WAKE) {
                live = P.state == PS_RESUMING;
            } else {
                live = P.state == PS_UNSTARTED;
            }
        }
        if (live) {
#if defined(ITSB_HIDDEN_FAST)
            if (P.prog == PROG_HIDDEN_WAIT_ONE) hidden_wait_one(H, idx); else
#endif
            {
            if (!is_hidden_prog(P.prog)) count_event(R, kind, P.prog, P.node, 0);
            runnable = idx;
            }
        }
    } else if (kind == ITSB_EV_TX_BEGIN) {
        on_tx_begin(H, ev_pkt(data));
    } else if (kind == ITSB_EV_TX_END) {
        if (on_tx_end(H, ev_pkt(data))) deliver = ev_pkt(data);
    } else if (kind == ITSB_EV_THROW) {
        Pcb& P = R.pcb[idx];
        if (P.state != PS_FREE && P.gen == ev_gen(data)) {
            if (!is_hidden_prog(P.prog) && !(ev_exc(data) & EXC_HIDDEN))
                count_event(R, ITSB_EV_THROW, P.prog, P.node, ev_exc(data));
            if (P.state == PS_UNSTARTED) proc_exit(H, idx);
            else runnable = deliver_exception(H, idx, ev_exc(data));
#if defined(ITSB_HIDDEN_FAST)
            /* a helper whose handler is the program's EXIT (CleanUp / CancelBalk: `except ...: pass`) ends here */
            if (runnable != NIL32 && is_hidden_prog(P.prog) && (u32)(R.m.code[P.pc] & 0xFF) == OP_EXIT) {
                proc_exit(H, idx);
                runnable = NIL32;
            }
#endif
        }
    } else {
        fail(H, ITSB_ERR_BAD_OPCODE, data);
    }
    if (deliver != NIL32) on_deliver(R, deliver);
    if (runnable != NIL32) run_process(R, runnable);
}

/* The key replicas are grouped by: the kind of the pending event.  Measured on the default bench (k_run_vote, 768
 * threads per SM): kind 3.72, (kind, program of the activated process) 3.30, (kind, program, pc) 3.32, three classes
 * (activation / transmission start / end + delivery) 3.01, two classes 2.66 x 10^9 events/s -- finer keys converge
 * better inside the handler but leave fewer lanes per round.  ITSB_KEY_WITH_PROGRAM selects the second form. */
ITSB_FN u32 pending_key(Rep& R, const PendingEvent& E) {
    const u32 kind = ev_kind(E.data);
#if defined(ITSB_KEY_WITH_PROGRAM)
    u32 prog = 0;
    if (kind == ITSB_EV_WAKE || kind == ITSB_EV_TIMER || kind == ITSB_EV_PROC_START || kind == ITSB_EV_THROW)
        prog = R.pcb[ev_idx(E.data)].prog & 15u;
    return (kind << 4) | prog;      /* < 256 */
#else
    (void)R;
    return kind << 4;               /* < 256 */
#endif
}

/* ---- connection records (overrides.py:102-120,152-186) ------------------------------------------------------------------
 * The override-path Socket files every packet it sends under its destination Location and every packet it receives
 * under its source Location, and logs a record when the two halves of a pair meet, when an unanswered outbound
 * packet is superseded, and for whatever is left when the socket closes.  The reference's two dictionaries never
 * forget (a daemon socket's grows by one entry per broadcast it hears); here each is a window of the newest
 * `connection_window` entries in insertion order, one 32-byte sector per entry in HBM, searched by the lanes in one
 * coalesced read.  oracle/netsimple.py applies the same window. */
ITSB_FN ConnEnt* conn_table(Hdr* H, u32 s, u32 dir) {
    return H->conntab + ((size_t)s * 2 + dir) * H->L.conn_window;
}
/* the replica's record log is conn_cap records followed by conn_cap order keys (u32): a record's key is the counter
 * value of the event whose processing logged it.  Records are appended in processing order except those of wake-ups
 * retired in the delivery pass, which are logged early under the key their wake-up would have had;
 * itsb_read_connections sorts by (End_Time, key), which is the order Socket.log_pair was called in. */
ITSB_FN u32* conn_order(Hdr* H) { return (u32*)(H->connlog + H->L.conn_cap); }
ITSB_FN u16* conn_fill(Hdr* H, u32 s) { return (u16*)((u8*)H + H->L.off_sockx) + 2 * s; }

ITSB_COLD int conn_find(const ConnEnt* T, u32 n, u32 ip, u32 port) {
    u32 hit = NIL32;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (u32 i = (u32)lane_id(); i < n; i += ITSB_W)
        if (T[i].key_ip == ip && (T[i].ports & 0xFFFF) == port) hit = i;
    const u32 m = w_min(hit);
    return m == NIL32 ? -1 : (int)m;
}

ITSB_COLD void conn_remove(ConnEnt* T, u16* fill, u32 idx) {
    const u32 n = *fill;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (u32 base = idx + 1; base < n; base += ITSB_W) {
        const u32 j = base + (u32)lane_id();
        ConnEnt E;
        if (j < n) E = T[j];
        w_sync();
        if (j < n) T[j - 1] = E;
        w_sync();
    }
    w_sync();
    *fill = (u16)(n - 1);
}

ITSB_COLD void conn_log(Hdr* H, u32 node, u32 has, u32 in_ip, u32 in_port, u32 in_size, u32 remote_ip, u32 out_port,
                      u32 out_size, double start, bool inbound);
ITSB_FN double conn_time(const ConnEnt& E) { return bitsd(((u64)E.t_hi << 32) | E.t_lo); }

/* what flush_logs writes for an entry that never became a pair (overrides.py:181-186) */
ITSB_FN void conn_log_single(Hdr* H, u32 node, const ConnEnt& E, u32 dir) {
    if (dir == 0) conn_log(H, node, 2, 0, 0, 0, E.a_ip, E.ports >> 16, E.size, conn_time(E), false);
    else conn_log(H, node, 1, E.a_ip, E.ports >> 16, E.size, 0, 0, 0, conn_time(E), true);
}

/* files an entry at position idx (>= 0: the key is already there and keeps its place, like a dict assignment) or
 * appends it; a full window first gives up its oldest entry, logged as flush_logs would have logged it at close */
ITSB_COLD void conn_put(Hdr* H, u32 node, ConnEnt* T, u16* fill, u32 dir, int idx, u32 key_ip, u32 a_ip, u32 size,
                      u32 ports) {
    if (idx < 0) {
        if (*fill >= H->L.conn_window) {
            const ConnEnt Old = T[0];
            conn_log_single(H, node, Old, dir);
            conn_remove(T, fill, 0);
        }
        idx = *fill;
        w_sync();
        *fill = (u16)(idx + 1);
    }
    if (lane_id() == 0) {
        ConnEnt E;
        const u64 tb = dbits(H->now);
        E.key_ip = key_ip; E.a_ip = a_ip; E.size = size; E.ports = ports;
        E.t_lo = (u32)tb; E.t_hi = (u32)(tb >> 32); E.pad[0] = 0; E.pad[1] = 0;
        T[idx] = E;
    }
    w_sync();
}

/* Socket.log_pair: in_* describe the inbound packet (has & 1), out_* the outbound one (has & 2) */
ITSB_COLD void conn_log(Hdr* H, u32 node, u32 has, u32 in_ip, u32 in_port, u32 in_size, u32 remote_ip, u32 out_port,
                      u32 out_size, double start, bool inbound) {
    { Rep Rs; Rs.stats = (u64*)((unsigned char*)H + H->L.off_stats); stat_add(Rs, ITSB_TM_USER0 + 6, 1); }   /* records produced (user slot 6) */
    const u32 n = H->conn_count;
    if (n < H->L.conn_cap && lane_id() == 0) {
        itsb_connection C;
        C.t_end = H->now; C.t_start = start;
        C.local_ip = (has & 1) ? in_ip : 0; C.local_port = (u16)((has & 1) ? in_port : 0);
        C.received_bytes = (has & 1) ? in_size : 0;
        C.remote_ip = (has & 2) ? remote_ip : 0; C.remote_port = (u16)((has & 2) ? out_port : 0);
        C.sent_bytes = (has & 2) ? out_size : 0;
        C.node = (u16)node; C.inbound = inbound ? 1 : 0; C.has = (u8)has;
        H->connlog[n] = C;
        conn_order(H)[n] = H->cur_seq;
    }
    H->conn_count = n + 1;
}

/* correlate_outbound_packet (overrides.py:102-109), called by send() after the packet went to the network */
ITSB_COLD void conn_outbound(Hdr* H, u32 s, u32 node, u32 dst_ip, u32 dst_port, u32 size, u32 tag, u32 f1) {
    const u32 remote = (tag < 32 && ((H->L.address_tags >> tag) & 1)) ? f1 : dst_ip;
    u16* fill = conn_fill(H, s);
    ConnEnt* TO = conn_table(H, s, 0);
    ConnEnt* TI = conn_table(H, s, 1);
    const int i = conn_find(TI, fill[1], dst_ip, dst_port);
    if (i >= 0) {
        const ConnEnt E = TI[i];
        conn_log(H, node, 3, E.a_ip, E.ports >> 16, E.size, remote, dst_port, size, conn_time(E), true);
        conn_remove(TI, fill + 1, (u32)i);
        return;
    }
    const int j = conn_find(TO, fill[0], dst_ip, dst_port);
    if (j >= 0) {
        const ConnEnt E = TO[j];
        conn_log(H, node, 2, 0, 0, 0, E.a_ip, E.ports >> 16, E.size, conn_time(E), false);
    }
    conn_put(H, node, TO, fill, 0, j, dst_ip, remote, size, dst_port | (dst_port << 16));
}

/* correlate_inbound_packet (overrides.py:115-120), called by recv() on the packet it returns */
ITSB_COLD void conn_inbound_fields(Hdr* H, u32 s, u32 src_ip, u32 ports, u32 dst_ip, u32 size);
ITSB_COLD void conn_inbound(Hdr* H, u32 s) {
    Rep R; rebind(R, H);
    const ArenaNode A = sock_front(R, R.sock[s]);
    conn_inbound_fields(H, s, A.src_ip, A.ports, A.pad, A.size);
}

/* The correlations of the wake-ups retired in one delivery pass (see on_deliver): what conn_inbound does, with one
 * lane per retired socket instead of the warp per socket -- the sockets are different, so are their tables, and the
 * chains of dependent HBM accesses (find, log, remove, put) overlap instead of queueing up.  Each socket logs at most
 * one record (the completed pair, or the entry its full window gives up); log slots go by lane rank = recipient order. */
ITSB_COLD void conn_retired(Hdr* H, u32 mask, u32 s_lane, u32 seq_lane, u32 k) {
    Rep R; rebind(R, H);
    const Pkt K = *pkt_at(R, k);
    const bool mine = ((mask >> lane_id()) & 1u) != 0;
    const u32 src_ip = K.src_ip, src_port = K.ports & 0xFFFF, dst_ip = K.dst_ip, dst_port = K.ports >> 16, size = K.size;
    const u32 window = H->L.conn_window;
    const u64 nowb = dbits(H->now);
    itsb_connection C;
    bool emit = false;
    if (mine) {
        u16* fill = conn_fill(H, s_lane);
        ConnEnt* TO = conn_table(H, s_lane, 0);
        ConnEnt* TI = conn_table(H, s_lane, 1);
        C.t_end = H->now; C.node = (u16)R.sock[s_lane].node;
        int j = -1;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (u32 i = 0; i < fill[0]; ++i)
            if (TO[i].key_ip == src_ip && (TO[i].ports & 0xFFFF) == src_port) { j = (int)i; break; }
        if (j >= 0) {                       /* the answer to something this socket sent: log the pair, forget the entry */
            const ConnEnt E = TO[j];
            C.t_start = conn_time(E);
            C.local_ip = dst_ip; C.local_port = (u16)dst_port; C.received_bytes = size;
            C.remote_ip = E.a_ip; C.remote_port = (u16)(E.ports >> 16); C.sent_bytes = E.size;
            C.inbound = 0; C.has = 3;
            emit = true;
            const u32 n = fill[0];
#if defined(__CUDACC__)
#pragma unroll 1
#endif
            for (u32 i = (u32)j + 1; i < n; ++i) TO[i - 1] = TO[i];
            fill[0] = (u16)(n - 1);
        } else {
            int i = -1;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
            for (u32 q = 0; q < fill[1]; ++q)
                if (TI[q].key_ip == src_ip && (TI[q].ports & 0xFFFF) == src_port) { i = (int)q; break; }
            if (i < 0) {
                u32 n = fill[1];
                if (n >= window) {          /* full window: its oldest entry is logged as flush_logs would log it */
                    const ConnEnt E = TI[0];
                    C.t_start = conn_time(E);
                    C.local_ip = E.a_ip; C.local_port = (u16)(E.ports >> 16); C.received_bytes = E.size;
                    C.remote_ip = 0; C.remote_port = 0; C.sent_bytes = 0;
                    C.inbound = 1; C.has = 1;
                    emit = true;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
                    for (u32 q = 1; q < n; ++q) TI[q - 1] = TI[q];
                    n -= 1;
                }
                i = (int)n;
                fill[1] = (u16)(n + 1);
            }
            ConnEnt E;
            E.key_ip = src_ip; E.a_ip = dst_ip; E.size = size; E.ports = src_port | (dst_port << 16);
            E.t_lo = (u32)nowb; E.t_hi = (u32)(nowb >> 32); E.pad[0] = 0; E.pad[1] = 0;
            TI[i] = E;
        }
    }
    const u32 emask = w_ballot(emit);
    const u32 total = (u32)popc(emask);
    const u32 n0 = H->conn_count;
    if (emit) {
        const u32 slot = n0 + (u32)popc(emask & lanemask_lt());
        if (slot < H->L.conn_cap) {
            H->connlog[slot] = C;
            conn_order(H)[slot] = seq_lane;
        }
    }
    w_sync();
    if (total) stat_add(R, ITSB_TM_USER0 + 6, total);      /* records produced (user slot 6) */
    H->conn_count = n0 + total;
    w_sync();
}

ITSB_COLD void conn_inbound_fields(Hdr* H, u32 s, u32 src_ip, u32 ports, u32 dst_ip, u32 size) {
    const u32 src_port = ports & 0xFFFF, dst_port = ports >> 16;
    u16* fill = conn_fill(H, s);
    ConnEnt* TO = conn_table(H, s, 0);
    ConnEnt* TI = conn_table(H, s, 1);
    const u32 node = ((Sock*)((unsigned char*)H + H->L.off_sock))[s].node;
    const int j = conn_find(TO, fill[0], src_ip, src_port);
    if (j >= 0) {
        const ConnEnt E = TO[j];
        conn_log(H, node, 3, dst_ip, dst_port, size, E.a_ip, E.ports >> 16, E.size, conn_time(E), false);
        conn_remove(TO, fill, (u32)j);
        return;
    }
    const int i = conn_find(TI, fill[1], src_ip, src_port);
    conn_put(H, node, TI, fill + 1, 1, i, src_ip, dst_ip, size, src_port | (dst_port << 16));
}

/* flush_logs (overrides.py:181-186): what never became a pair, when the socket closes */
ITSB_COLD void conn_flush(Hdr* H, u32 s, u32 node) {
    u16* fill = conn_fill(H, s);
    ConnEnt* TO = conn_table(H, s, 0);
    ConnEnt* TI = conn_table(H, s, 1);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (u32 i = 0; i < fill[0]; ++i) {
        const ConnEnt E = TO[i];
        conn_log_single(H, node, E, 0);
    }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (u32 i = 0; i < fill[1]; ++i) {
        const ConnEnt E = TI[i];
        conn_log_single(H, node, E, 1);
    }
    w_sync();
    fill[0] = 0; fill[1] = 0;
}

/* The instructions a model that keeps connection records is assembled with (Asm.log_connections): the socket calls
 * of the override path followed by the correlation the reference's Socket does inside them.  Kept apart from
 * exec_cold so that the code every other model runs is the same with or without this feature. */
ITSB_COLD int exec_logged(Hdr* H, u32 p, u64 ins) {
    Rep R; rebind(R, H);
    Pcb& P = R.pcb[p];
    const u32 op = (u32)(ins & 0xFF), a = (u32)((ins >> 8) & 0xFF), b = (u32)((ins >> 16) & 0xFF),
              c = (u32)((ins >> 24) & 0xFF), imm = (u32)(ins >> 32);
    switch (op) {
    case OP_RECV_LOG: {   /* OP_RECV (run_process), then correlate_inbound_packet on the packet returned */
        if (P.sock == NIL16) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return XR_STOP; }
        Sock& S = R.sock[P.sock];
        if (S.q_len == 0) {
            waiters_append(R, S.wait_head, S.wait_tail, p, (u16)(WS_SOCK | P.sock));
            P.state = PS_WAIT;
            node_digest_waits(R, S.node, P.sock);
            return XR_STOP;
        }
        if (H->conntab) conn_inbound(H, P.sock);
        const ArenaNode A = sock_pop(R, S);
        set_pkt_regs(R, A);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_RECV_FILTER_LOG: {   /* OP_RECV_FILTER (run_process) of a model that keeps connection records: every packet
                                    the loop head pops is correlated first, handed to the program or not; what a
                                    filtered wake-up would do is then fixed, so the delivery pass can retire it */
        if (P.sock == NIL16) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return XR_STOP; }
        NodeDyn& N = R.node[P.node];
        Sock& S = R.sock[P.sock];
        const itsb_filter F = R.m.filters[b];
        for (;;) {
            if (P.phase == 0) {               /* with ws.awake(): wait_until_awake, remember the wake epoch */
                if (!N.awake) {
                    waiters_append(R, N.wake_head, N.wake_tail, p, (u16)(WS_NODE | P.node));
                    P.state = PS_WAIT;
                    return XR_STOP;
                }
                P.r[a & 3] = N.epoch;
                P.phase = 1;
            }
            if (S.q_len == 0) {               /* socket.recv() */
                waiters_append(R, S.wait_head, S.wait_tail, p, (u16)(WS_SOCK | P.sock));
                P.state = PS_WAIT;
                node_digest_parked(R, S.node, P.sock, p, b, a);
                return XR_STOP;
            }
            if (H->conntab) conn_inbound(H, P.sock);
            const ArenaNode A = sock_pop(R, S);
            P.phase = 0;
            if (N.epoch != P.r[a & 3]) { P.pc = (u16)(imm & 0xFFFF); return XR_CONTINUE; }      /* FellAsleep */
            if (filter_passes(R, F, A.tag, A.f0, R.m.nodes[P.node].host, P)) {
                set_pkt_regs(R, A);
                P.pc += 1;
                return XR_CONTINUE;
            }
        }
    }
    case OP_OPEN_AT:      /* a second socket, on another attachment of this machine: imm = node | port << 16 */
        if (sock_open_on(H, p, imm & 0xFFFF, imm >> 16, false) < 0) return XR_STOP;
        P.pc += 1;
        return XR_CONTINUE;
    case OP_SEND_LOG: {   /* OP_SEND, then correlate_outbound_packet */
        const u32* v = H->vol;
        u32 dst_ip, dst_port = imm & 0xFFFF;
        switch (a) {
        case DST_BCAST: dst_ip = R.m.nets[R.m.nodes[P.node].net].bcast; break;
        case DST_PKT_SRC: dst_ip = v[SRC_PKT_SRC_IP - 0x10]; dst_port = v[SRC_PKT_SRC_PORT - 0x10]; break;
        case DST_ENT: dst_ip = v[SRC_ENT_IP - 0x10]; dst_port = v[SRC_ENT_PORT - 0x10]; break;
        case DST_SELF: dst_ip = R.m.nodes[P.node].addr; break;
        default: dst_ip = P.r[2]; dst_port = P.r[3] & 0xFFFF; break;
        }
        u32 f0s = (imm >> 16) & 0xFF, f1s = (imm >> 24) & 0xFF;
        if (P.sock == NIL16) { fail(H, ITSB_EXC_SOCKET_CLOSED, P.pc); return XR_STOP; }
        const u32 size = src_val(R, P, b), f1 = f1s == SRC_NONE ? 0 : src_val(R, P, f1s);
        net_send(H, P.node, R.sock[P.sock].port, dst_ip, dst_port, size, c, f0s == SRC_NONE ? 0 : src_val(R, P, f0s), f1);
        if (H->conntab && H->status == 0) conn_outbound(H, P.sock, P.node, dst_ip, dst_port, size, c, f1);
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_FORWARD_LOG: {   /* OP_FORWARD, then the correlation on the socket of the other attachment */
        const u32* v = H->vol;
        net_send(H, imm & 0xFFFF, imm >> 16, P.r[2], P.r[3] & 0xFFFF, v[SRC_PKT_SIZE - 0x10], v[SRC_PKT_TAG - 0x10],
                 v[SRC_PKT_F0 - 0x10], v[SRC_PKT_F1 - 0x10]);
        if (H->conntab && H->status == 0) {
            u16 ws = R.node[imm & 0xFFFF].sock_head;
            while (ws != NIL16 && R.sock[ws].port != (imm >> 16)) ws = R.sock[ws].next;
            if (ws != NIL16)
                conn_outbound(H, ws, imm & 0xFFFF, P.r[2], P.r[3] & 0xFFFF, v[SRC_PKT_SIZE - 0x10],
                              v[SRC_PKT_TAG - 0x10], v[SRC_PKT_F1 - 0x10]);
        }
        P.pc += 1;
        return XR_CONTINUE;
    }
    case OP_STAT_ADD:
        stat_add(R, ITSB_TM_USER(imm), src_val(R, P, a));
        P.pc += 1;
        return XR_CONTINUE;
    case OP_STAT_FIRST:
        if (stat_get(R, ITSB_TM_USER(imm)) == 0) stat_add(R, ITSB_TM_USER(imm), (u64)(H->now * 1.0e6));
        P.pc += 1;
        return XR_CONTINUE;
    case OP_CLOSE_LOG:    /* OP_CLOSE after flush_logs (overrides.py:366-367) */
        if (P.sock != NIL16 && H->conntab) conn_flush(H, P.sock, R.sock[P.sock].node);
        sock_close(H, p);
        P.pc += 1;
        return XR_CONTINUE;
    default:
        fail(H, ITSB_ERR_BAD_OPCODE, ins);
        return XR_STOP;
    }
}

/* Builds the t = 0 state of a replica: empty tables, free lists, and the model's installed processes in
 * installation order (Endpoint.install -> sim.add, overrides.py:423-424; network_simple.py:366-385).  The header's
 * binding fields must already be staged (bind_views). */
ITSB_FN void init_replica(Rep& R, u64 replica_id) {
    Hdr* H = R.hdr;
    const Layout* L = &H->L;
    H->now = 0.0; H->draws = 0; H->seq = 0; H->status = 0; H->detail = 0;
    H->trace_count = 0; H->log_count = 0; H->conn_count = 0;
    H->replica_lo = (u32)replica_id; H->replica_hi = (u32)(replica_id >> 32);
    H->nq_head = 0; H->nq_tail = 0; H->nq_spill_head = NIL32; H->nq_spill_tail = NIL32; H->nq_spill_count = 0;
    H->ev_count = 0; H->spare32 = 0;
    H->ovf_min_t = T_EMPTY; H->ovf_min_seq = NIL32; H->ovf_head = NIL32; H->ovf_count = 0;
    H->spare16 = 0; H->cur_seq = 0; H->rng_c0 = 0; H->rng_c1 = 0;
    for (u32 i = 0; i < VOL_COUNT; ++i) H->vol[i] = 0;
    for (u32 i = 0; i < 2 * NQ_CAP; ++i) R.nq[i] = 0;
    for (u32 i = 0; i < L->max_ev; ++i) {
        R.ev_heap[i].t = T_EMPTY; R.ev_heap[i].seq = 0; R.ev_heap[i].slot = 0;
        R.ev_pos[i] = 0; R.ev_data[i] = 0; R.ev_free[i] = (u16)(L->max_ev - 1 - i);
    }
    H->ev_free_top = (u16)L->max_ev;
    for (u32 i = 0; i < L->max_proc; ++i) { memset(&R.pcb[i], 0, sizeof(Pcb)); R.pcb_free[i] = (u16)(L->max_proc - 1 - i); }
    H->pcb_free_top = (u16)L->max_proc;
    for (u32 i = 0; i < L->max_sock; ++i) {
        memset(&R.sock[i], 0, sizeof(Sock)); memset(&R.sockpkt[i], 0, sizeof(ArenaNode));
        R.sockv[i].count = 0; R.sockv[i].bytes = 0;
        R.sock_free[i] = (u16)(L->max_sock - 1 - i);
    }
    H->sock_free_top = (u16)L->max_sock;
    for (u32 i = 0; i < L->n_nodes; ++i) {
        NodeDyn& N = R.node[i];
        N.sock_head = NIL16; N.port_cursor = 1024; N.wake_head = NIL16; N.wake_tail = NIL16; N.epoch = 0; N.awake = 0;
    }
    for (u32 i = 0; i < L->n_nodes; ++i) {   /* no socket yet */
        R.ndig[i] = L->max_sig != 0 ? 0u : ((u32)ND_SINGLE | nd_sock_field(ND_NOSOCK));
        R.nmask[i] = 0;
    }
    for (u32 i = 0; i < L->max_proc; ++i) R.pmask[i] = 0;
    for (u32 i = 0; i < ITSB_TM_SIZE; ++i) R.stats[i] = 0;
    H->arena_free_top = L->arena_nodes;
    H->sig_free_top = (u16)L->max_sig; H->thr_free_top = (u16)L->max_thr; H->task_free_top = (u16)L->max_task;
    H->setup_done = 0;
    H->hidden_entry_balk = 0; H->hidden_entry_wait_one = 0;
    for (u32 i = 0; i < R.m.n_entries; ++i) {
        if (R.m.entries[i].prog == PROG_HIDDEN_BALK) H->hidden_entry_balk = i;
        if (R.m.entries[i].prog == PROG_HIDDEN_WAIT_ONE) H->hidden_entry_wait_one = i;
    }
    for (u32 i = 0; i < L->n_ifaces; ++i) R.iface_addr[i] = R.m.ifaces[i].addr0;
    for (u32 i = 0; i < L->table_words; ++i) R.tab[i] = 0;
    for (u32 i = 0; i < L->max_sig; ++i) { memset(&R.sig[i], 0, sizeof(Sig)); R.sig_free[i] = (u16)(L->max_sig - 1 - i); }
    for (u32 i = 0; i < L->max_thr; ++i) { memset(&R.thr[i], 0, sizeof(Thr)); R.thr_free[i] = (u16)(L->max_thr - 1 - i); }
    for (u32 i = 0; i < L->max_task; ++i) { memset(&R.task[i], 0, sizeof(Task)); R.task_free[i] = (u16)(L->max_task - 1 - i); }
    if (L->max_sig != 0) {
        for (u32 i = 0; i < L->max_proc; ++i) memset(&R.pcbx[i], 0, sizeof(PcbX));
        for (u32 i = 0; i < L->n_nodes; ++i) {
            if ((R.m.nodes[i].flags >> 16) != 0) { R.node[i].port_cursor = 32768; R.node[i].awake = 0; }
        }
    }
    for (u32 i = 0; i < R.m.n_init; ++i) {
        const itsb_initproc I = R.m.init[i];
        if (I.mode == ITSB_INIT_RUN_PROC) m_run_proc_in(H, I.node, NIL32, I.entry, I.delay_const, I.r);
        else proc_spawn(H, I.entry, I.node, I.r, 0.0);
    }
}

}  // namespace itsb
