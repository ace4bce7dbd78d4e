/* sdr_pipe_t32c.cu -- the receiver pipeline kernel for 32-sample tiles, CONTRACTING build (see sdr_pipe_tu.cuh).
 * Opt-in per handle (sdr_batch_desc.flags & SDR_BATCH_CONTRACT): fused multiply-adds and history-first summation in the linear
 * sections -- biquad cascades, Hilbert FIR, NCO complex multiply; nothing next to a compare or a truncation.  Results stay
 * within north_star's 1e-4 of full scale but are not bit-identical to the reference; the default build is the exact one. */
#define SDR_FIXED_T 32
#define SDR_TSUF _t32c
#define SDR_NS sdrk32c
#define SDR_LB_THREADS 448
#define SDR_LB_BLOCKS 1
#define SDR_CONTRACT
#ifndef SDR_HANDOVER
#define SDR_LOCKSTEP
#endif
#include "sdr_pipe_tu.cuh"
