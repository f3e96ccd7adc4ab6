/*
 * itsb_pool.cu -- k_run_pool for models whose tables are staged in shared memory (the flagship path); every measured
 * shape is instantiated here (ITSB_POOL_SHAPE picks one, default 2 = 512 threads x 1 CTA per SM, 768 slots).
 */
#define POOL_ALL_SHAPES 1
#include "itsb_pool.cuh"
