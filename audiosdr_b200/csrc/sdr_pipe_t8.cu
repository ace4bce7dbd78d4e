/* sdr_pipe_t8.cu -- the receiver pipeline kernel for 8-sample tiles (see sdr_pipe_tu.cuh). */
#define SDR_FIXED_T 8
#define SDR_TSUF _t8
#define SDR_NS sdrk8
#define SDR_LB_THREADS 352
#define SDR_LB_BLOCKS 2
/* shorter tiles (plans that share an SM): mbarrier hand-over between the stages; -DSDR_LEAN_LOCKSTEP builds the lock-step form */
#ifdef SDR_LEAN_LOCKSTEP
#define SDR_LOCKSTEP
#endif
#include "sdr_pipe_tu.cuh"
