/*
 * itsb.h -- C ABI of libitsim_b200.so, the B200-native batched discrete-event engine behind ITsim's modelling API.
 *
 * The reference has no FFI: its seam is Python-level (ITsim objects calling greensim).  Each
 * entry point below names the reference interface it replaces (paths relative to /root/reference).  Plain pointers
 * and sizes only; the caller (numpy / ctypes) owns every buffer passed in or out, the library
 * copies on load and owns all device memory until itsb_destroy.  One host thread drives one engine; engines on
 * different GPUs are independent.  There is no CPU execution path: without a CUDA device every compute entry
 * point fails with ITSB_ERR_NO_DEVICE.
 *
 * Return convention: 0 = OK; negative = engine error (ITSB_ERR_*); positive = a model-level exception of the
 * reference raised in some replica (ITSB_EXC_*), retrievable per replica with itsb_replica_status.
 */
#ifndef ITSB_H
#define ITSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ITSB_ABI_VERSION 5

/* ---- engine errors (negative) ------------------------------------------------------------------------------ */
#define ITSB_OK                    0
#define ITSB_ERR_NO_DEVICE        -1   /* no CUDA device / driver: the product has no CPU fallback            */
#define ITSB_ERR_BAD_MODEL        -2   /* malformed tables (bounds, sizes, versions)                          */
#define ITSB_ERR_CUDA             -3   /* a CUDA runtime call failed; see itsb_last_error                     */
#define ITSB_ERR_STATE            -4   /* call order (e.g. run before init_replicas)                          */
#define ITSB_ERR_CAP_EVENTS       -10  /* per-replica event pool exhausted                                    */
#define ITSB_ERR_CAP_PROCS        -11  /* per-replica process slots exhausted                                 */
#define ITSB_ERR_CAP_SOCKETS      -12  /* per-replica socket slots exhausted                                  */
#define ITSB_ERR_CAP_PACKETS      -13  /* reserved                                                            */
#define ITSB_ERR_CAP_ARENA        -14  /* per-replica queue arena (HBM) exhausted                             */
#define ITSB_ERR_BAD_OPCODE       -15
#define ITSB_ERR_CAP_THREADS      -16  /* Thread / Process / Signal table exhausted                             */
#define ITSB_ERR_NCCL             -20

/* ---- model-level exceptions of the reference (positive) ------------------------------------------------------ */
#define ITSB_EXC_NO_SUCH_ADDRESS   1   /* itsim/network/link.py:18-22; overrides CannotForward :293-297        */
#define ITSB_EXC_NO_ROUTE          2   /* itsim/machine/node.py:59-66                                          */
#define ITSB_EXC_PORT_IN_USE       3   /* itsim/machine/node.py:41-49; network_simple_overrides.py:343-344     */
#define ITSB_EXC_PORTS_EXHAUSTED   4   /* itsim/machine/node.py:52-56                                          */
#define ITSB_EXC_SOCKET_CLOSED     5   /* ValueError("Socket is closed"), itsim/machine/socket.py:105,130,138  */
#define ITSB_EXC_NEGATIVE_DELAY    6   /* greensim Simulator._schedule ValueError                              */
#define ITSB_EXC_UNCAUGHT          7   /* a non-Interrupt exception left a process body (aborts Simulator.run) */

/* ---- trace / event kinds: the shared definition of a "simulated event"  -------------- */
#define ITSB_EV_PROC_START 1
#define ITSB_EV_TIMER      2
#define ITSB_EV_TX_BEGIN   3
#define ITSB_EV_TX_END     4
#define ITSB_EV_DELIVER    5
#define ITSB_EV_WAKE       6
#define ITSB_EV_THROW      7
#define ITSB_EV_STOP       8
#define ITSB_NUM_KINDS     9
#define ITSB_TRACE_MARK    9   /* trace-only record (not an event): a model side effect declared with the MARK op */

typedef struct itsb_trace_rec {   /* 16 bytes, little endian; same layout as oracle/evtrace.py TRACE_DTYPE */
    double   t;
    uint8_t  kind;
    uint8_t  prog;
    uint16_t node;
    uint32_t aux;
} itsb_trace_rec;

typedef struct itsb_net_event {   /* 32 bytes: one record of the datastore's `network_event` table
                                     (itsim/schemas/items.py:83-131): a socket opened or closed                     */
    double   t;                   /* simulated time                                                                */
    uint16_t node;
    uint16_t port;
    uint8_t  type;                /* 1 = "open", 2 = "close"                                                       */
    uint8_t  protocol;            /* 0 NONE, 1 UDP, 2 TCP                                                          */
    int16_t  pid;
    uint32_t src_ip;              /* local address the socket is bound on                                          */
    uint32_t dst_ip;              /* 0: datagram sockets have no peer at open time                                 */
    uint16_t dst_port;
    uint16_t prog;                /* program id of the process that opened / closed it                             */
    uint16_t tags;                /* its tags (bit i = Tag value i: 0 MALWARE, 1 VULNERABLE)                      */
    uint16_t reserved;
} itsb_net_event;

typedef struct itsb_connection {  /* 40 bytes: one connection record, the JSON `Socket.log_pair` prints
                                     (examples/network_simple/network_simple_overrides.py:152-177)                  */
    double   t_end;               /* End_Time: when the pair was logged                                            */
    double   t_start;             /* Start_Time: when the first packet of the pair was sent / received             */
    uint32_t local_ip;            /* Local_IP: destination of the inbound packet (0: none)                         */
    uint32_t remote_ip;           /* Remote_IP: ADDRESS field of the outbound payload, else its destination (0)    */
    uint32_t sent_bytes;          /* Sent_Bytes: size of the outbound packet (0: none)                             */
    uint32_t received_bytes;      /* Received_Bytes: size of the inbound packet (0: none)                          */
    uint16_t local_port, remote_port;
    uint16_t node;
    uint8_t  inbound;             /* Direction: 1 "Inbound", 0 "Outbound"                                          */
    uint8_t  has;                 /* bit 0: an inbound packet is part of the pair, bit 1: an outbound one          */
} itsb_connection;

/* ---- telemetry: what the datastore would have been told, reduced on device ---------------------------------- */
#define ITSB_LAT_BINS 64
/* slots of the uint64 telemetry vector returned by itsb_read_telemetry / reduced by itsb_allreduce_telemetry */
#define ITSB_TM_KIND0            0                      /* [0..8]  events per kind (index = ITSB_EV_*)           */
#define ITSB_TM_PACKETS_SENT     9
#define ITSB_TM_BYTES_SENT       10
#define ITSB_TM_PACKETS_RECEIVED 11                     /* packets returned by a recv()                        */
#define ITSB_TM_BYTES_RECEIVED   12
#define ITSB_TM_DELIVERED        13                     /* deliveries that found a socket                      */
#define ITSB_TM_DROPPED          14                     /* deliveries dropped at the node                      */
#define ITSB_TM_USER0            15                     /* [15..22] model counters (STAT op); slot 22 counts the
                                                           network_event records, slot 21 the connection
                                                           records, when those logs are on                     */
#define ITSB_TM_LAT0             23                     /* [23..86] transit-delay histogram, log2-spaced bins  */
#define ITSB_TM_USER8            (ITSB_TM_LAT0 + ITSB_LAT_BINS)   /* [87..90] model counters 8..11              */
#define ITSB_TM_USER_SLOTS       12
#define ITSB_TM_SIZE             (ITSB_TM_USER8 + 4)
#define ITSB_TM_USER(s)          ((s) < 8 ? ITSB_TM_USER0 + (s) : ITSB_TM_USER8 + (s) - 8)

/* ---- model description: flat struct-of-arrays tables produced by the host-side compiler ---------------------- */

typedef struct itsb_dist {        /* one variate generator chain (itsim/random.py:14-29 + greensim.random)       */
    uint32_t kind;                /* ITSB_DIST_*                                                                 */
    uint32_t flags;               /* ITSB_DF_*                                                                   */
    double   p0, p1;              /* constant: p0; uniform: a,b; expo: lambd=1/mean; normal: mu,sigma; choice: n */
    double   slope, shift;        /* linear(gen, slope, shift)            (ITSB_DF_LINEAR)                       */
    double   lower, upper;        /* bounded(gen, lower, upper)           (ITSB_DF_LOWER / ITSB_DF_UPPER)        */
} itsb_dist;

#define ITSB_DIST_CONSTANT 0
#define ITSB_DIST_UNIFORM  1
#define ITSB_DIST_EXPO     2
#define ITSB_DIST_NORMAL   3
#define ITSB_DIST_CHOICE   4
#define ITSB_DF_LINEAR 1u
#define ITSB_DF_LOWER  2u
#define ITSB_DF_UPPER  4u
#define ITSB_DF_INT    8u         /* project_int: truncate toward zero                                           */

typedef struct itsb_net {         /* one broadcast domain: override Network (overrides.py:189-291) or Link        */
    uint32_t cidr_net, cidr_mask; /* network address and netmask                                                 */
    uint32_t bcast;               /* broadcast address                                                           */
    uint32_t first_node, n_nodes; /* ITSB_NET_OVERRIDE: members = node indices [first_node, first_node + n_nodes), in
                                     link order; ITSB_NET_LINK: members = members[first_node .. first_node + n_nodes),
                                     in connection order                                                          */
    uint32_t lat_dist, bw_dist;   /* indices into the distribution table (already bounded as the reference does)  */
    uint32_t mode;                /* ITSB_NET_OVERRIDE: TX process, delay = lat + len/bw ; ITSB_NET_LINK: draw at
                                     send, delay = lat + 8*len/bw (itsim/network/link.py:82-86)                  */
} itsb_net;
#define ITSB_NET_OVERRIDE 0
#define ITSB_NET_LINK     1

typedef struct itsb_node {        /* static part of a node                                                        */
    uint32_t addr;                /* IPv4 address on its network                                                  */
    uint32_t net;                 /* index into nets                                                              */
    uint32_t host;                /* override nodes: hostname id (payload field f0 compares against it);
                                     shipped-itsim nodes: ITSB_NODE_* bits                                        */
    uint32_t flags;               /* shipped-itsim nodes: first interface | number of interfaces << 16            */
} itsb_node;
#define ITSB_NODE_FORWARDS 1u     /* a Router whose handle_packet_transit forwards:
                                     interface, hop = _solve_transfer(dest); interface.link._transfer_packet(...)  */

typedef struct itsb_iface {       /* Interface(link, address, routes) (itsim/network/interface.py:22-68); a node's
                                     interfaces are consecutive, loopback first (itsim/machine/node.py:74-77)      */
    uint32_t node, net;
    uint32_t addr0;               /* address at t = 0 (DHCP may change it at run time)                            */
    uint32_t first_route, n_routes; /* [Local(link.cidr)] + routes, in the order _solve_transfer scans them        */
    uint32_t reserved[3];
} itsb_iface;

typedef struct itsb_route {       /* Local / Relay (itsim/network/route.py:8-59)                                  */
    uint32_t net, mask, prefixlen;
    uint32_t kind;                /* ITSB_ROUTE_LOCAL: hop = destination; ITSB_ROUTE_RELAY: hop = gateway          */
    uint32_t gateway;
    uint32_t reserved[3];
} itsb_route;
#define ITSB_ROUTE_LOCAL 0
#define ITSB_ROUTE_RELAY 1

typedef struct itsb_entry {       /* program entry point                                                          */
    uint32_t pc;                  /* first instruction                                                            */
    uint16_t prog;                /* program id reported in trace records                                         */
    uint16_t tags;                /* greensim tags of the body (`@tagged`, itsim/__init__.py:9-11,52-57): bit i =
                                     Tag value i.  A process carries its body's tags and those of the process that
                                     created it (greensim cascades them)                                          */
} itsb_entry;

typedef struct itsb_initproc {    /* process started at t=0, in program order                                     */
    uint32_t entry, node;
    uint32_t r[4];
    uint32_t mode;                /* ITSB_INIT_ADD: sim.add(body) (Endpoint.install, overrides.py:423-424);
                                     ITSB_INIT_RUN_PROC: Node.run_proc_in(sim, delay, body) (node.py:279-283)      */
    uint32_t delay_const;         /* index of the delay in consts (ITSB_INIT_RUN_PROC)                            */
} itsb_initproc;
#define ITSB_INIT_ADD      0
#define ITSB_INIT_RUN_PROC 1
#define ITSB_INIT_ADD_DRAWN 2  /* sim.add on node `node + next(dists[delay_const])`, drawn per replica at set-up
                                  (infected = next(distribution(workstations)), malware_simple.py:114)            */
#define ITSB_INIT_ADD_SAME  3  /* sim.add on the node the previous ITSB_INIT_ADD_DRAWN entry drew                 */

typedef struct itsb_filter {      /* which received packets a RECV_FILTER loop hands to its program; the rest it
                                     consumes and loops (the daemon bodies' "not for me" branches,
                                     network_simple.py:187-221,234-245).  Bit i of a mask = payload tag i.          */
    uint32_t tags_always;         /* always handed over                                                           */
    uint32_t tags_if_host;        /* handed over only if the packet's hostname field equals the node's             */
    uint32_t tags_if_list;        /* handed over only if the process remembers at least one packet (LIST_*)        */
    uint32_t tags_if_listed;      /* handed over only if the process remembers a packet with the same hostname
                                     field (the entry a LIST_TAKE on that field would find)                        */
} itsb_filter;

typedef struct itsb_caps {        /* per-replica capacities; exceeding one flags the replica with ITSB_ERR_CAP_*  */
    uint32_t max_events, max_procs, max_sockets;
    uint32_t max_packets;         /* reserved: packets in flight are arena nodes                                  */
    uint32_t arena_nodes;         /* 32-byte nodes in HBM: socket queues, remembered packets, packets in flight   */
    uint32_t trace_records;       /* per-replica trace capacity (0 = tracing off)                                 */
    uint32_t max_threads, max_tasks, max_signals;   /* shipped-itsim Thread / Process / Signal tables (0 = none)  */
    uint32_t table_words;         /* per-replica uint32 model table (TLOAD / TSTORE), zero-initialised            */
    uint32_t net_event_records;   /* per-replica capacity of the network_event log in HBM (0 = logging off)       */
    uint32_t connection_records;  /* per-replica capacity of the connection log in HBM (0 = correlation off)      */
    uint32_t connection_window;   /* packets a socket remembers per direction while waiting for the other half of
                                     a pair (1..4096; the reference's dictionaries are unbounded, overrides.py:80-81;
                                     a full window logs its oldest entry as unanswered, as flush_logs would at
                                     close)                                                                       */
} itsb_caps;

typedef struct itsb_model_desc {
    uint32_t abi_version;
    uint32_t n_code, n_entries, n_dists, n_consts, n_nets, n_nodes, n_init, n_filters;
    uint32_t n_ifaces, n_routes, n_members;
    uint32_t address_tags;        /* bit i: payloads tagged i carry an ADDRESS entry in field f1 (Remote_IP of the
                                     connection records, overrides.py:158-162)                                    */
    itsb_caps caps;
} itsb_model_desc;

/* blobs[] order for itsb_load_model */
#define ITSB_BLOB_CODE    0   /* uint64[n_code]      instruction words (see itsim_b200/ir.py)                      */
#define ITSB_BLOB_ENTRIES 1   /* itsb_entry[n_entries]                                                            */
#define ITSB_BLOB_DISTS   2   /* itsb_dist[n_dists]                                                               */
#define ITSB_BLOB_CONSTS  3   /* double[n_consts]                                                                 */
#define ITSB_BLOB_NETS    4   /* itsb_net[n_nets]                                                                 */
#define ITSB_BLOB_NODES   5   /* itsb_node[n_nodes]                                                               */
#define ITSB_BLOB_INIT    6   /* itsb_initproc[n_init]                                                            */
#define ITSB_BLOB_FILTERS 7   /* itsb_filter[n_filters]                                                           */
#define ITSB_BLOB_IFACES  8   /* itsb_iface[n_ifaces]                                                             */
#define ITSB_BLOB_ROUTES  9   /* itsb_route[n_routes]                                                             */
#define ITSB_BLOB_MEMBERS 10  /* uint32[n_members] node indices of ITSB_NET_LINK networks                         */
#define ITSB_NUM_BLOBS    11

typedef struct itsb_engine itsb_engine;

/* Replaces: Simulator() construction (itsim/simulator.py:9; greensim.Simulator.__init__).  Binds a CUDA device. */
int itsb_create(int device, itsb_engine** out);

/* Replaces: the Python object graph a model script builds (Link/Network, Endpoint, install(): network_simple.py:
 * 331-387; itsim/machine/node.py:84-133,279-283).  Copies the tables to the device.                              */
int itsb_load_model(itsb_engine* e, const itsb_model_desc* desc, const void* const* blobs, const size_t* sizes);

/* Replaces: launching N interpreters with N seeds.  Replica r uses Philox key (seed, first_replica_id + r); all
 * replicas start from the model's initial state at t = 0.  Allocates per-replica state in HBM.                   */
int itsb_init_replicas(itsb_engine* e, uint64_t n, uint64_t seed, uint64_t first_replica_id);

/* Replaces: Simulator.run(duration) (greensim; subclassed at itsim/simulator.py:9) for every replica at once.
 * Re-entrant like the reference: state stays on the device, a later call continues.  events_done (may be NULL)
 * receives the number of simulated events (section 8-d definition) processed by this call over all replicas.
 * Returns 0, or the first non-zero replica status (itsb_replica_status tells which).                              */
int itsb_run(itsb_engine* e, double duration, uint64_t* events_done);

/* Asynchronous halves of itsb_run, for callers that overlap host work or time with their own events.            */
int itsb_run_async(itsb_engine* e, double duration);
int itsb_sync(itsb_engine* e, uint64_t* events_done);
/* Device time of the last completed run in milliseconds (CUDA events on the engine's stream).                   */
int itsb_last_run_ms(itsb_engine* e, float* ms);
/* Number of kernel launches issued by this engine so far.                                                        */
int itsb_launch_count(itsb_engine* e, uint64_t* launches);

/* Name of the engine kernel itsb_run launches for the loaded model ("k_run_vote", "k_run", "k_run_in_place", ...; ""
 * before a model is loaded).  No reference counterpart: the reference has one interpreter loop.                  */
const char* itsb_kernel_name(itsb_engine* e);

/* Replaces: Simulator.now() per replica.  dst receives n doubles starting at replica `first`.                     */
int itsb_read_now(itsb_engine* e, uint64_t first, uint64_t n, double* dst);

/* Replaces: reading the event log of one replica (the reference has none; the oracle's tracer defines it).       */
int itsb_read_trace(itsb_engine* e, uint64_t replica, itsb_trace_rec* dst, size_t cap, size_t* n_total);

/* Replaces: the `network_event` records the datastore stores (itsim/schemas/items.py:83-131; itsim/datastore/
 * database.py:44-57).  dst receives up to cap records of one replica; n_total = records produced (may exceed the
 * capacity the model asked for: later ones are counted but not kept).                                               */
int itsb_read_net_events(itsb_engine* e, uint64_t replica, itsb_net_event* dst, size_t cap, size_t* n_total);

/* Histogram over the replicas of one telemetry slot (the per-replica statistics of BASELINE.json config 3: beacons
 * per replica, bytes exfiltrated, time of the first beacon): dst[b] = number of replicas whose value v falls in bin
 * b = min((v - lo) / width, nbins - 1), values below lo in bin 0; with `only_nonzero` replicas whose value is 0 are
 * left out (a time that was never set).  comm (an ncclComm_t, may be NULL) sums the bins over the ranks first.  */
int itsb_histogram(itsb_engine* e, uint32_t slot, uint64_t lo, uint64_t width, uint32_t nbins, int only_nonzero,
                   void* comm, uint64_t* dst);

/* The connection records of one replica, in the order `Socket.log_pair` was called
 * (network_simple_overrides.py:102-120,152-186).  *n_total = records produced (may exceed the capacity kept).
 * The device appends a record when it is produced and keeps, next to it, the counter value of the event that
 * produced it; records of wake-ups retired in the delivery pass are produced early, under the value their wake-up
 * would have had, and this call orders what the log holds by (End_Time, that value) before copying it out.     */
int itsb_read_connections(itsb_engine* e, uint64_t replica, itsb_connection* dst, size_t cap, size_t* n_total);

/* ---- streaming the record logs home -------------------------------------------------------------------------------
 * Replaces: the record-at-a-time path of the reference -- logging.Handler.emit -> one HTTP POST -> one SQLite connection
 * per record (itsim/logging.py:27-65, itsim/datastore/database.py:130-164) and the per-connection JSON lines of
 * Socket.log_pair (network_simple_overrides.py:152-186).  The per-replica logs in HBM are fixed-capacity windows; a
 * drain empties all of them at once:
 *   itsb_drain_begin  (engine stream) counts every replica's new records, packs them replica after replica into a
 *                     device staging buffer with ONE gather kernel -- connection records put in the order log_pair was
 *                     called in (End_Time, then the counter value of the producing event), as itsb_read_connections
 *                     does -- resets the per-replica counts, and starts ONE cudaMemcpyAsync per log on a second stream
 *                     into pinned host memory.  The next itsb_run may be queued right away: it overlaps the copy.
 *   itsb_drain_end    waits for that copy and hands out pointers into the pinned buffers.  Two sets of buffers
 *                     alternate, so the pointers stay valid until the drain after the next one begins; at most one
 *                     drain may be pending besides the one being begun.
 * Record r's replica: net_first / conn_first hold n_replicas + 1 offsets -- replica i owns [first[i], first[i + 1]).   */
typedef struct itsb_drain {
    const itsb_net_event*  net;        uint64_t n_net;
    const uint64_t*        net_first;
    const itsb_connection* conn;       uint64_t n_conn;
    const uint64_t*        conn_first;
    uint64_t net_dropped, conn_dropped;   /* produced since the last drain beyond a replica's capacity: counted, lost */
    uint64_t d2h_bytes;                   /* bytes this drain copied device -> host                                   */
    double   pack_ms;                     /* device time of the count / scan / gather kernels (CUDA events)          */
} itsb_drain;
int itsb_drain_begin(itsb_engine* e);
int itsb_drain_end(itsb_engine* e, itsb_drain* out);

/* Replaces: the records itsim/logging.py + itsim/datastore would receive (README.md:18-50), reduced on the device:
 * ITSB_TM_SIZE uint64 counters / histogram bins summed over this engine's replicas.                               */
int itsb_read_telemetry(itsb_engine* e, uint64_t* dst, size_t n);

/* Per-replica counters (ITSB_TM_SIZE uint64 each), for parity checks on sampled replicas.                          */
int itsb_read_replica_telemetry(itsb_engine* e, uint64_t replica, uint64_t* dst, size_t n);

/* Sums the telemetry vector over all ranks (ncclAllReduce, SUM, uint64) -- the only collective on this path.
 * nccl_comm is an ncclComm_t created by the caller; the reduced vector is left on the device and in dst.          */
int itsb_allreduce_telemetry(itsb_engine* e, void* nccl_comm, uint64_t* dst, size_t n);
/* uint64 all-reduce of a small host vector over the ranks (op: 0 = sum, 2 = max, ncclRedOp_t): what a one-process-per-
 * GPU host needs besides the telemetry sum -- a barrier, the slowest rank's time, the job's event count -- without
 * a second communication library.  buf (n <= 4096) is read, reduced on the device through NCCL, and written back. */
int itsb_allreduce_u64(itsb_engine* e, void* nccl_comm, uint64_t* buf, size_t n, int op);
/* Helpers so a ctypes host can build the communicator without an NCCL binding of its own.                         */
int itsb_nccl_unique_id(void* id128, size_t cap);
int itsb_nccl_comm_init(itsb_engine* e, int nranks, int rank, const void* id128, void** comm_out);
int itsb_nccl_comm_destroy(void* comm);

/* status of one replica: 0, ITSB_ERR_CAP_* or ITSB_EXC_*; detail = offending value (address, port, pc...).        */
int itsb_replica_status(itsb_engine* e, uint64_t replica, int32_t* status, uint64_t* detail);
/* index of the first replica with non-zero status, or -1.                                                         */
int64_t itsb_first_failed_replica(itsb_engine* e);

/* bytes of per-replica state staged HBM<->shared memory per run (S_in + S_out of the roofline accounting).        */
int itsb_state_bytes(itsb_engine* e, uint64_t* hot_bytes, uint64_t* arena_bytes);

/* Replaces: math.log inside random.expovariate / random.normalvariate (CPython Lib/random.py; reached through
 * greensim.random.expo / normal: itsim/random.py:4, itsim/network/service/dhcp/client.py:47-49).  Evaluates the
 * engine's own logarithm -- the one its variates use -- on n doubles ON THE DEVICE (one thread per value), so a
 * parity test can hold it bit for bit against the oracle's restatement of the same algorithm.                     */
int itsb_eval_log(itsb_engine* e, const double* x, double* y, size_t n);

const char* itsb_last_error(itsb_engine* e);
void itsb_destroy(itsb_engine* e);
int itsb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ITSB_H */
