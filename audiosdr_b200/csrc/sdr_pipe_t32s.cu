/* sdr_pipe_t32s.cu -- the receiver pipeline kernel for 32-sample tiles, SSB-class buckets only (see sdr_pipe_tu.cuh).
 * The same code as sdr_pipe_t32.cu with the pipeline class as a compile-time constant: ring offsets that depend on the class
 * are literals instead of selects and the ENV-class stages (PLL, envelope path) are not in the kernel, which the 14 warps'
 * instruction streams share the instruction caches with.  SSB buckets of the default (exact) build launch this kernel; ENV
 * buckets on 32-sample tiles the general one. */
#define SDR_FIXED_T 32
#define SDR_FIXED_CLS 0 /* CLS_SSB */
#define SDR_TSUF _t32s
#define SDR_NS sdrk32s
#define SDR_LB_THREADS 448
#define SDR_LB_BLOCKS 1
#ifndef SDR_HANDOVER
#define SDR_LOCKSTEP
#endif
#include "sdr_pipe_tu.cuh"
