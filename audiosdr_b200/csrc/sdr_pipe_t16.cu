/* sdr_pipe_t16.cu -- the receiver pipeline kernel for 16-sample tiles (see sdr_pipe_tu.cuh). */
#define SDR_FIXED_T 16
#define SDR_TSUF _t16
#define SDR_NS sdrk16
#define SDR_LB_THREADS 352
#define SDR_LB_BLOCKS 2
/* The stages run the lock-step schedule of the hand-over rules (one CTA-wide barrier per tile step): measured faster than the
 * mbarrier hand-over on every workload and tile length; -DSDR_HANDOVER builds the other form. */
#ifndef SDR_HANDOVER
#define SDR_LOCKSTEP
#endif
#include "sdr_pipe_tu.cuh"
