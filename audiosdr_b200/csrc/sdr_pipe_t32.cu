/* sdr_pipe_t32.cu -- the receiver pipeline kernel for 32-sample tiles (see sdr_pipe_tu.cuh). */
#define SDR_FIXED_T 32
#define SDR_TSUF _t32
#define SDR_NS sdrk32
#define SDR_LB_THREADS 448
#define SDR_LB_BLOCKS 1
#include "sdr_pipe_tu.cuh"
