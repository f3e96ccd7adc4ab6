/*
 * itsb_wide.cu -- the engine kernel for models whose replica fits shared memory 13 to 16 times per SM: same code
 * as k_run, compiled for 512 threads per CTA (128 registers per thread).  k_run itself is compiled for the 12 warps
 * the flagship model runs with, which leaves it 168 registers and no spills.  Its own translation unit: a second
 * kernel in k_run's cubin moves the hot loop's code placement (measured).
 */
#include "itsb_kernel.cuh"

extern "C" __global__ void __launch_bounds__(512, 1) k_run_wide(const RunParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    run_warps<false>(p, smem);
}
