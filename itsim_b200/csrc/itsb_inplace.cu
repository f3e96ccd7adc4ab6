/*
 * itsb_inplace.cu -- the engine kernel for replicas too large for shared memory (config C4: the 1 040-node LAN is
 * ~650 KB per replica): same warp-per-replica code, state advanced in place in HBM, cached by L1/L2 (a replica is
 * touched by one warp of one SM during a launch).  16-20 warps per SM instead of the 12 that shared memory allows.
 */
#define ITSB_HIDDEN_FAST 1      /* select()'s helper hops without the interpreter (itsb_core.cuh: hidden_wait_one) */
#include "itsb_kernel.cuh"

extern "C" __global__ void __launch_bounds__(640, 1) k_run_in_place(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    run_warps<true>(p, smem);
}
