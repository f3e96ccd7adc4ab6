/*
 * itsb_pool_hbm.cu -- k_run_pool for models whose tables do not fit shared memory (config C4), read in place, with
 * greensim's select() helper hops taken without the interpreter (ITSB_HIDDEN_FAST), like k_run_in_place and
 * k_run_vote_hbm.  One shape (512 x 1 / 768).
 */
#define ITSB_HIDDEN_FAST 1
#define POOL_ENTRY(name) name##_hbm
#include "itsb_pool.cuh"
