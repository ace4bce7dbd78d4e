/* sdr_pipe_t32.cu -- the receiver pipeline kernel for 32-sample tiles (see sdr_pipe_tu.cuh). */
#define SDR_FIXED_T 32
#define SDR_TSUF _t32
#define SDR_NS sdrk32
#define SDR_LB_THREADS 448
#define SDR_LB_BLOCKS 1
/* The stages run the lock-step schedule of the hand-over rules (one CTA-wide barrier per tile step): measured faster than the
 * mbarrier hand-over on every workload and tile length; -DSDR_HANDOVER builds the other form. */
#ifndef SDR_HANDOVER
#define SDR_LOCKSTEP
#endif
#include "sdr_pipe_tu.cuh"
