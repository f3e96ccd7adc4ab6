/*
 * itsb_lanes_hbm.cu -- k_run_vote_hbm: the lane-per-replica kernel for models whose tables do not fit shared memory
 * (config C4: 1 024 endpoints + 16 forwarding routers, 234 KB of tables, 667 KB of state per replica).  Same body as
 * k_run_vote (itsb_vote.cuh); the tables are read where they lie (they are read-only: L1/L2 keep them), and greensim's
 * select() helper hops -- five of the seven heap pops behind a packet received on the shipped-itsim path -- run without
 * the interpreter (ITSB_HIDDEN_FAST, see hidden_wait_one in itsb_core.cuh).
 *
 * Why a lane and not a block per replica: on this path nothing in the engine is lane-parallel (Link._transfer_packet's
 * recipients are walked one after the other, m_deliver in itsb_machine.cuh), so a warp per replica -- k_run_in_place --
 * spends 32 lanes on one replica's scalar chain and holds 2 368 replicas in flight per GPU, while a lane per replica
 * holds all 10 000 of config C4 in flight, spread three to a warp (RunParams.lanes_per_warp).
 */
#define ITSB_HIDDEN_FAST 1
#include "itsb_vote.cuh"

extern "C" __global__ void __launch_bounds__(256, ITSB_LANES_MIN_CTAS) k_run_vote_hbm(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    run_vote(p, smem);
}
