/* sdr_kernel.h -- launcher interface between sdr_host.cpp and the kernels (sdr_kernel.cu). */
#ifndef SDR_KERNEL_H
#define SDR_KERNEL_H
#include <stdint.h>
#include "sdr_types.h"

/* state-reset bits: which reference setter side effect to replay on the device state */
enum {
  SDRK_R_IF = 1u,  /* arm_biquad_cascade_df1_init_f32 on both IF rails   (setDemodMode, C:187-222) */
  SDRK_R_IMG = 2u, /* ... on the AM image rails                           (init, C:178-179)        */
  SDRK_R_AUD = 4u, /* ... on the audio filter                             (setAudioFilter, C:298-311) */
  SDRK_R_ALS = 8u, /* taps and ring zeroed                                (enableALSfilter, C:384-391) */
  SDRK_R_NB = 16u  /* initBlanker                                         (C:676-682)              */
};

#ifdef __cplusplus
extern "C" {
#endif
int sdrk_setup_device(const float *hilbert64);
int sdrk_launch_pipeline(const SdrLaunch *L, void *stream);
int sdrk_launch_als_pass(const SdrLaunch *L, void *stream); /* second launch of a split ALS bucket (L->lay from lay_build_als, L->raw) */
int sdrk_occupancy(const SdrLaunch *L); /* resident CTAs per SM granted to this launch configuration (diagnostics) */
int sdrk_launch_reset(float *state, unsigned long long ch_stride, const uint32_t *chan, const uint32_t *mask, uint32_t n, void *stream);
int sdrk_launch_fill_word(float *state, unsigned long long ch_stride, uint32_t w, float v, uint32_t n_ch, void *stream);
int sdrk_launch_gather(const float *state, unsigned long long ch_stride, const uint32_t *chan, uint32_t n, const uint32_t *words,
                       uint32_t n_words, float *out, void *stream); /* words == NULL: all SDR_STATE_WORDS in order */
int sdrk_launch_scatter(float *state, unsigned long long ch_stride, const uint32_t *chan, uint32_t n, const float *in, void *stream);
#ifdef __cplusplus
}
#endif
#endif
