/* oracle/sdr_aux_oracle.c -- TEST INFRASTRUCTURE (the checker), not product code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's CPU legs may load this library.
 *
 * Plain-C restatement of the two blocks either side of the receiver chain:
 *   ora_pp_*  AudioSDRpreProcessor::update()   /root/reference/SRC/AudioSDRlib/AudioSDRpreProcessor.cpp:46-138  ("PP")
 *   ora_iq_*  AudioIQgenerator::update()       /root/reference/SRC/AudioSDRlib/AudioIQgenerator.cpp:33-87       ("IQ")
 * PINNED against the unmodified reference compiled on the host (oracle/_ref/refaux, tests/test_aux_oracle.py) and the
 * committed fixtures tests/golden/aux_*.npz -- with ONE exception stated in oracle/aux_fft128.h: the pre-processor's FFT
 * is CMSIS-DSP code that is not in the reference tree, so both this file and refaux use the restated transform
 * ("parity unpinned" for FFT rounding only; the integer data path and the decision logic are pinned).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "aux_fft128.h"

#define NB 128

typedef struct { uint32_t channel, block, opcode; float a0; } aux_event;

/* ------------------------------------------------------------------ pre-processor ---- */
typedef struct {
  int16_t I2Scorrection, savedSample, failureCount, successCount; /* PP.h:66-69 */
  int IQswap, autoDetectFlag;                                      /* PP.h:70,72 */
  float buffer[256];                                               /* PP.h:63 */
} pp_t;

static void pp_apply(pp_t *p, const aux_event *e) {
  switch (e->opcode) {
    case 1: p->autoDetectFlag = 1; p->I2Scorrection = 0; p->failureCount = 0; p->successCount = 0; break; /* PP.cpp:142-148 */
    case 2: p->autoDetectFlag = 0; p->I2Scorrection = 0; break;                                           /* PP.cpp:151-154 */
    case 3: p->I2Scorrection = (int16_t)(int)e->a0; p->autoDetectFlag = 0; break;                         /* PP.cpp:160-163 */
    case 4: p->IQswap = e->a0 != 0.0f; break;                                                             /* PP.cpp:169 */
    default: break;
  }
}

static void pp_update(pp_t *p, int16_t *I, int16_t *Q) {
  /* one-sample lag compensation, PP.cpp:61-71.  The -1 branch shifts Q but writes the saved sample into I[0]: as written. */
  if (p->I2Scorrection == 1) {
    int16_t temp = I[NB - 1];
    for (int i = NB - 1; i > 0; i--) I[i] = I[i - 1];
    I[0] = p->savedSample;
    p->savedSample = temp;
  } else if (p->I2Scorrection == -1) {
    int16_t temp = Q[NB - 1];
    for (int i = NB - 1; i > 0; i--) Q[i] = Q[i - 1];
    I[0] = p->savedSample;
    p->savedSample = temp;
  }
  /* image detector, PP.cpp:81-118 */
  if (p->autoDetectFlag) {
    const int n_FFT = 128, min = 5;
    int maxLine = 0;
    float *buffer = p->buffer;
    for (int i = 0; i < 128; i++) {
      buffer[2 * i] = (float)((double)(float)I[i] / 32767.0);
      buffer[2 * i + 1] = (float)((double)(float)Q[i] / 32767.0);
    }
    aux_cfft128_forward(buffer);
    aux_cmplx_mag_squared(buffer, buffer, 128);
    float average_power = 0.0f, maximum_power = 0.0f;
    for (int i = min; i < n_FFT - min; i++) {
      average_power += buffer[i];
      if (buffer[i] > maximum_power) { maxLine = i; maximum_power = buffer[i]; }
    }
    average_power /= (float)(n_FFT - 2 * min);
    float imbalance_ratio = maximum_power / buffer[n_FFT - maxLine];
    if ((double)maximum_power > 10.0 * (double)average_power) {
      if ((double)imbalance_ratio < 10.0) p->failureCount++;
      else p->failureCount = 0;
      if (p->failureCount > 10) {
        p->I2Scorrection++;
        if (p->I2Scorrection > 1) p->I2Scorrection = -1;
        p->failureCount = 0;
        p->successCount = 0;
      }
      p->successCount++;
    }
    if (p->successCount > 1000) p->autoDetectFlag = 0;
  }
  /* I/Q swap, PP.cpp:124-130 */
  if (p->IQswap) {
    for (int i = 0; i < 128; i++) { int16_t t = I[i]; I[i] = Q[i]; Q[i] = t; }
  }
}

/* ------------------------------------------------------------------ I/Q generator ---- */
typedef struct {
  float bufferI[3 * NB], bufferQ[3 * NB]; /* function statics, IQ.cpp:37-38 */
  float gainI, gainQ;                     /* IQ.h:72-73 */
} iq_t;

static float iq_tap(int k) { float f; memcpy(&f, &AUX_IQ_HILBERT[k], 4); return f; }

static void iq_apply(iq_t *g, const aux_event *e) {
  if (e->opcode == 1) { g->gainI = e->a0; g->gainQ = (float)(1.0 / (double)e->a0); } /* IQ.h:56-60 */
}

/* (int16_t)(double): what the host-compiled reference does (cvttsd2si to 32 bits, then the low half) */
static int16_t to_i16(double d) {
  int32_t t = (d >= 2147483648.0 || d <= -2147483649.0 || d != d) ? (int32_t)0x80000000 : (int32_t)d;
  return (int16_t)t;
}

static void iq_update(iq_t *g, const int16_t *x, int16_t *I, int16_t *Q) {
  const int L = 257, D = 128;
  float Idata[NB], Qdata[NB];
  for (int i = 0; i < NB; i++) { /* IQ.cpp:52-60 */
    const float v = (float)((double)(float)x[i] / 32767.0);
    g->bufferI[i] = g->bufferI[NB + i]; g->bufferI[NB + i] = g->bufferI[2 * NB + i]; g->bufferI[2 * NB + i] = v;
    g->bufferQ[i] = g->bufferQ[NB + i]; g->bufferQ[NB + i] = g->bufferQ[2 * NB + i]; g->bufferQ[2 * NB + i] = v;
  }
  for (int i = 0; i < NB; i++) { /* IQ.cpp:65-76 */
    float acc = 0.0f;
    for (int k = 0; k < L / 4; k++) {
      const int i1 = (2 * NB + i) - (2 * k + 1), i2 = (2 * NB + i) - L + 2 * (k + 1);
      const float d = g->bufferQ[i1] - g->bufferQ[i2];
      const float p = iq_tap(k) * d;
      acc = acc + p;
    }
    Qdata[i] = acc;
    Idata[i] = g->bufferI[2 * NB + i - D];
  }
  for (int i = 0; i < NB; i++) { /* IQ.cpp:78-82 */
    I[i] = to_i16((double)Idata[i] * 32767.0 * (double)g->gainI);
    Q[i] = to_i16((double)Qdata[i] * 32767.0 * (double)g->gainQ);
  }
}

/* ------------------------------------------------------------------ batch runners ---- */
typedef struct {
  int kind; uint32_t c0, c1, n_blocks, n_events; const aux_event *ev;
  const int16_t *in0, *in1; int16_t *out0, *out1; int32_t *status;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  const size_t ns = (size_t)j->n_blocks * NB;
  for (uint32_t c = j->c0; c < j->c1; c++) {
    if (j->kind == 1) {
      pp_t p; memset(&p, 0, sizeof p);
      int16_t I[NB], Q[NB];
      for (uint32_t b = 0; b <= j->n_blocks; b++) {
        for (uint32_t e = 0; e < j->n_events; e++) /* stable order within a block; trailing events after the last block */
          if ((j->ev[e].channel == c || j->ev[e].channel == 0xFFFFFFFFu) &&
              (b < j->n_blocks ? j->ev[e].block == b : j->ev[e].block >= b)) pp_apply(&p, &j->ev[e]);
        if (b == j->n_blocks) break;
        memcpy(I, j->in0 + c * ns + (size_t)b * NB, 2 * NB);
        memcpy(Q, j->in1 + c * ns + (size_t)b * NB, 2 * NB);
        pp_update(&p, I, Q);
        memcpy(j->out0 + c * ns + (size_t)b * NB, I, 2 * NB);
        memcpy(j->out1 + c * ns + (size_t)b * NB, Q, 2 * NB);
      }
      if (j->status) {
        int32_t *s = j->status + (size_t)c * 8;
        s[0] = p.autoDetectFlag; s[1] = p.I2Scorrection; s[2] = p.failureCount; s[3] = p.successCount;
        s[4] = p.savedSample; s[5] = p.IQswap; s[6] = s[7] = 0;
      }
    } else {
      iq_t g; memset(&g, 0, sizeof g); g.gainI = g.gainQ = 1.0f;
      for (uint32_t b = 0; b < j->n_blocks; b++) {
        for (uint32_t e = 0; e < j->n_events; e++)
          if ((j->ev[e].channel == c || j->ev[e].channel == 0xFFFFFFFFu) && j->ev[e].block == b) iq_apply(&g, &j->ev[e]);
        iq_update(&g, j->in0 + c * ns + (size_t)b * NB, j->out0 + c * ns + (size_t)b * NB, j->out1 + c * ns + (size_t)b * NB);
      }
      if (j->status) memset(j->status + (size_t)c * 8, 0, 32);
    }
  }
  return 0;
}

/* kind 1 = pre-processor (in0 = I, in1 = Q), kind 2 = generator (in0 = X, in1 unused).  Events must be sorted by block.
 * status: int32 [n_channels][8] or NULL.  Returns 0. */
int ora_aux_run(int kind, uint32_t n_channels, uint32_t n_blocks, const void *events, uint32_t n_events,
                const int16_t *in0, const int16_t *in1, int16_t *out0, int16_t *out1, int32_t *status, int threads) {
  if (threads < 1) threads = 1;
  if ((uint32_t)threads > n_channels) threads = (int)(n_channels ? n_channels : 1);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * threads);
  for (int t = 0; t < threads; t++) {
    job_t j = {kind, (uint32_t)((uint64_t)n_channels * t / threads), (uint32_t)((uint64_t)n_channels * (t + 1) / threads),
               n_blocks, n_events, (const aux_event *)events, in0, in1, out0, out1, status};
    jobs[t] = j;
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
  free(th); free(jobs);
  return 0;
}

/* ------------------------------------------------------------------ grabber ----
 * AudioGrabberComplex256::update() / grab(), AudioGrabberComplex256.cpp:50-91 ("GR"): no arithmetic.  Runs n_blocks updates
 * from a fresh object per channel, then one grab(): out[c][512] = the interleaved (re, im) samples of the last COMPLETE pair
 * of blocks (untouched = -1 in every slot when no pair has completed yet), flags[c] = {newDataAvailable() before the grab,
 * _dataBufferValid}. */
void ora_grab_run(uint32_t n_channels, uint32_t n_blocks, const int16_t *I, const int16_t *Q, int32_t *out, int32_t *flags) {
  const size_t ns = (size_t)n_blocks * NB;
  for (uint32_t c = 0; c < n_channels; c++) {
    int16_t buffer[512], outBuffer[512];
    uint16_t buffStart = 0;
    int valid = 0, fresh = 0;
    memset(buffer, 0, sizeof buffer); memset(outBuffer, 0, sizeof outBuffer);
    for (uint32_t b = 0; b < n_blocks; b++) { /* GR.cpp:58-69 (_transferringData is false outside grab()) */
      const int16_t *re = I + c * ns + (size_t)b * NB, *im = Q + c * ns + (size_t)b * NB;
      int16_t *dst = buffer + buffStart;
      for (int i = 0; i < NB; i++) { *dst++ = re[i]; *dst++ = im[i]; } /* GR.cpp:39-47 */
      buffStart = (uint16_t)((buffStart + 256) % 512);
      if (buffStart == 0) {
        for (int i = 0; i < 512; i++) outBuffer[i] = buffer[i];
        fresh = 1; valid = 1;
      }
    }
    flags[2 * c] = fresh; flags[2 * c + 1] = valid;
    for (int i = 0; i < 512; i++) out[(size_t)c * 512 + i] = valid ? outBuffer[i] : -1; /* GR.cpp:83-88 */
  }
}

/* Spectrum tap on grabber snapshots (the consumer of grab(): the sketch's panadapter, not library code -- SURVEY 8f row 3):
 * snap[c][512] interleaved (re, im) int16 -> power[c][256] = |FFT256|^2, samples taken as (float)int16, natural bin order.
 * PARITY UNPINNED against the sketch (its code is not in the reference tree; CMSIS arm_cfft_f32 is restated, see aux_fft128.h);
 * pinned against the mathematical DFT in tests/test_aux_oracle.py and bit for bit against the CUDA kernel. */
void ora_grab_spectrum(uint32_t n_channels, const int16_t *snap, float *power) {
  for (uint32_t c = 0; c < n_channels; c++) {
    float buf[512];
    for (int i = 0; i < 512; i++) buf[i] = (float)snap[(size_t)c * 512 + i];
    aux_cfft256_forward(buf);
    aux_cmplx_mag_squared(buf, power + (size_t)c * 256, 256);
  }
}

/* natural-order power spectrum of one block as the detector sees it (for the float64-DFT decision test) */
void ora_aux_power128(const int16_t *I, const int16_t *Q, float *power) {
  float buf[256];
  for (int i = 0; i < 128; i++) {
    buf[2 * i] = (float)((double)(float)I[i] / 32767.0);
    buf[2 * i + 1] = (float)((double)(float)Q[i] / 32767.0);
  }
  aux_cfft128_forward(buf);
  aux_cmplx_mag_squared(buf, power, 128);
}
