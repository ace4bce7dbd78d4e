/* Test-infrastructure shim (NOT product code).
 * The reference's only CMSIS-DSP calls on the hot path are
 * arm_biquad_cascade_df1_init_f32 / arm_biquad_cascade_df1_f32
 * (declared in the vendored header "ARM_MATH UPDATE/.../arm_math.h":1257-1262,
 * 1360-1378; call sites AudioSDR.cpp:77,78,136,137,285 and 175-218,300-309).
 * CMSIS-DSP V1.4.5b ("arm_math Ver 4.5.0", 20 Oct 2015) ships only as Cortex-M4
 * static libraries in the reference, so its published algorithm is restated
 * here: direct-form-I cascade, per stage
 *     acc = (b0*x) + (b1*x1) + (b2*x2) + (a1*y1) + (a2*y2)
 * evaluated left to right in float32, state order {x1,x2,y1,y2} per stage,
 * coefficients {b0,b1,b2,a1,a2} per stage (feedback signs pre-negated), each
 * stage filtering the whole block before the next stage reads it.
 * init stores the pointers and zeroes 4*numStages state words. */
#ifndef ORACLE_SHIM_ARM_MATH_H
#define ORACLE_SHIM_ARM_MATH_H
#include <stdint.h>
#include <string.h>
typedef float float32_t;
typedef double float64_t;
typedef struct {
  uint32_t numStages;
  float32_t *pState;
  float32_t *pCoeffs;
} arm_biquad_casd_df1_inst_f32;

static inline void arm_biquad_cascade_df1_init_f32(arm_biquad_casd_df1_inst_f32 *S, uint8_t numStages,
                                                   float32_t *pCoeffs, float32_t *pState) {
  S->numStages = numStages;
  S->pCoeffs = pCoeffs;
  memset(pState, 0, (4u * (uint32_t)numStages) * sizeof(float32_t));
  S->pState = pState;
}

static inline void arm_biquad_cascade_df1_f32(const arm_biquad_casd_df1_inst_f32 *S, float32_t *pSrc,
                                              float32_t *pDst, uint32_t blockSize) {
  float32_t *pIn = pSrc;
  float32_t *pState = S->pState;
  const float32_t *pCoeffs = S->pCoeffs;
  for (uint32_t stage = 0; stage < S->numStages; stage++) {
    float32_t b0 = *pCoeffs++, b1 = *pCoeffs++, b2 = *pCoeffs++, a1 = *pCoeffs++, a2 = *pCoeffs++;
    float32_t Xn1 = pState[0], Xn2 = pState[1], Yn1 = pState[2], Yn2 = pState[3];
    float32_t *pOut = pDst;
    for (uint32_t n = 0; n < blockSize; n++) {
      float32_t Xn = pIn[n];
      float32_t acc = (b0 * Xn) + (b1 * Xn1) + (b2 * Xn2) + (a1 * Yn1) + (a2 * Yn2);
      pOut[n] = acc;
      Xn2 = Xn1; Xn1 = Xn; Yn2 = Yn1; Yn1 = acc;
    }
    pState[0] = Xn1; pState[1] = Xn2; pState[2] = Yn1; pState[3] = Yn2;
    pState += 4;
    pIn = pDst;
  }
}

/* --- FFT entry points used by AudioSDRpreProcessor.cpp:88-89.  CMSIS-DSP itself is not in the reference tree; the
 *     transform is the restatement in oracle/aux_fft128.h (see its header: rounding unpinned, decisions checked). */
#include "../aux_fft128.h"
typedef struct { uint16_t fftLen; } arm_cfft_instance_f32;
static const arm_cfft_instance_f32 arm_cfft_sR_f32_len128 = {128};
static inline void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag) {
  (void)ifftFlag; (void)bitReverseFlag; /* the reference calls (.., 0, 1): forward, natural-order output */
  if (S->fftLen == 128) aux_cfft128_forward(p1);
}
static inline void arm_cmplx_mag_squared_f32(float32_t *pSrc, float32_t *pDst, uint32_t numSamples) {
  aux_cmplx_mag_squared(pSrc, pDst, numSamples);
}
#endif
