/* sdr_aux_core.cuh -- per-lane arithmetic of the pre-processor and the I/Q generator kernels (sdr_aux.cu).
 *
 * Everything here is __host__ __device__ so that tests/emu/aux_emu.cpp can run the very same source on the host against
 * the oracle; the kernels in sdr_aux.cu add only the data movement around it.
 *
 *   pre-processor  AudioSDRpreProcessor::update()  AudioSDRpreProcessor.cpp:46-138  ("PP")
 *   generator      AudioIQgenerator::update()      AudioIQgenerator.cpp:33-87       ("IQ")
 *
 * Arithmetic rules as in the receiver kernel: one rounding per operation (nvcc -fmad=false), explicit fmaf only where
 * an exact residual is wanted, IEEE division.
 */
#ifndef SDR_AUX_CORE_CUH
#define SDR_AUX_CORE_CUH
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AUX_HD __host__ __device__ __forceinline__
#define AUX_UNROLL _Pragma("unroll")
#define AUX_UNROLLN(n) _Pragma(AUX_STR(unroll n))
#define AUX_STR(x) #x
#else
#define AUX_HD inline
#define AUX_UNROLL
#define AUX_UNROLLN(n)
#endif

/* ------------------------------------------------------------------ shared helpers ---- */

/* (float)((double)(float)q / 32767.0) for an int16 q (PP.cpp:84-85, IQ.cpp:56,59): the correctly rounded float quotient,
 * by one Markstein correction of q * fl(1/32767) (tests/emu/aux_emu.cpp checks all 65 536 values against the divide). */
AUX_HD float aux_q15_to_float(int q) {
  const float r = 1.0f / 32767.0f; /* folded at compile time, correctly rounded */
  const float n = (float)q;
  const float d0 = n * r;
  const float e = fmaf(-d0, 32767.0f, n);
  return fmaf(e, r, d0);
}

/* (int16_t)(v * 32767.0 * gain) as the host-compiled reference evaluates it (IQ.cpp:80-81): double products, truncation
 * to 32 bits (x86 cvttsd2si), low half kept.  gain == 1 (the constructor's value, IQ.h:72-73) needs no FP64: the product
 * v * 32767 is exact in double, so its truncation follows from the float product and its exact FMA residual. */
AUX_HD int aux_to_pcm(float v, float gain) {
  const float a = fabsf(v);
  if (gain == 1.0f && a < 256.0f) {
    const float hi = a * 32767.0f;
    const float err = fmaf(a, 32767.0f, -hi);
    int r = (int)hi; /* hi < 2^23: integer part exact */
    if ((float)r == hi && err < 0.0f) r -= 1;
    if (v < 0.0f) r = -r;
    return (int)(int16_t)r;
  }
  const double d = (double)v * 32767.0 * (double)gain;
  int i;
  if (d >= 2147483648.0 || d <= -2147483649.0 || d != d) i = (int)0x80000000;
  else i = (int)d;
  return (int)(int16_t)i;
}

/* the gain == 1 case of aux_to_pcm for |v| < 256, branch-free (the I/Q generator's outputs never exceed 2 * sum|h| < 5) */
AUX_HD int aux_to_pcm_unit(float v) {
  const float a = fabsf(v);
  const float hi = a * 32767.0f;
  const float err = fmaf(a, 32767.0f, -hi);
  int r = (int)hi;
  r -= ((float)r == hi && err < 0.0f) ? 1 : 0;
  return v < 0.0f ? -r : r;
}

AUX_HD int aux_brev7(int i) {
#if defined(__CUDA_ARCH__)
  return (int)(__brev((unsigned)i) >> 25);
#else
  int r = 0;
  for (int k = 0; k < 7; k++) r |= ((i >> k) & 1) << (6 - k);
  return r;
#endif
}

/* ------------------------------------------------------------------ pre-processor ---- */
struct PpState { /* PP.h:66-72; one per channel in device memory */
  int32_t corr;   /* I2Scorrection  */
  int32_t saved;  /* savedSample    */
  int32_t fail;   /* failureCount   */
  int32_t succ;   /* successCount   */
  int32_t swap;   /* IQswap         */
  int32_t autod;  /* autoDetectFlag */
  int32_t pad0, pad1;
};

struct PpSetterCall { uint32_t channel; /* 0xFFFFFFFF = all */ uint32_t setter; int32_t arg; uint32_t pad; };

AUX_HD void pp_apply(PpState &s, uint32_t setter, int32_t arg) {
  switch (setter) {
    case 1: s.autod = 1; s.corr = 0; s.fail = 0; s.succ = 0; break; /* PP.cpp:142-148 */
    case 2: s.autod = 0; s.corr = 0; break;                         /* PP.cpp:151-154 */
    case 3: s.corr = (int32_t)(int16_t)arg; s.autod = 0; break;     /* PP.cpp:160-163 */
    case 4: s.swap = arg != 0; break;                               /* PP.cpp:169 */
    default: break;
  }
}

/* Feed-forward form of the lag compensation + swap for a channel whose detector is off (its correction cannot change during
 * the call): outputs for one chunk of 8 samples (chunks never straddle a block), straight from the input planes.  PP.cpp:61-71,124-130.
 *   corr = +1: I is delayed by one sample across block boundaries (savedSample carries the last one).
 *   corr = -1: inside every block Q[i] <- Q[i-1] for i >= 1, Q[0] stays, and I[0] <- the previous block's last ORIGINAL Q
 *              sample (the source writes blockI->data[0] in this branch, PP.cpp:69): reproduced, not repaired. */
AUX_HD void pp_static_chunk(const int16_t *vi_in, const int16_t *vq_in, int16_t prev_i, int16_t prev_q, bool block_head, int corr,
                            int swap, int16_t *oi, int16_t *oq) {
  /* vi_in/vq_in: the 8 input samples of the chunk; prev_*: the input sample just before it (savedSample at the start of
   * the call); block_head: the chunk starts a 128-sample block */
  int16_t vi[8], vq[8];
  AUX_UNROLL for (int t = 0; t < 8; t++) { vi[t] = vi_in[t]; vq[t] = vq_in[t]; }
  if (corr == 1) {
    AUX_UNROLL for (int t = 7; t > 0; t--) vi[t] = vi[t - 1];
    vi[0] = prev_i;
  } else if (corr == -1) {
    const int16_t q0 = vq[0];
    AUX_UNROLL for (int t = 7; t > 0; t--) vq[t] = vq[t - 1];
    vq[0] = block_head ? q0 : prev_q;
    if (block_head) vi[0] = prev_q;
  }
  AUX_UNROLL for (int t = 0; t < 8; t++) { oi[t] = swap ? vq[t] : vi[t]; oq[t] = swap ? vi[t] : vq[t]; }
}

/* Sequential form on one block held in a row of `stride`-spaced int16 (the detector path): PP.cpp:61-71 as written. */
AUX_HD void pp_correct_block(int16_t *I, int16_t *Q, PpState &s) {
  if (s.corr == 1) {
    const int16_t temp = I[127];
    for (int i = 127; i > 0; i--) I[i] = I[i - 1];
    I[0] = (int16_t)s.saved;
    s.saved = temp;
  } else if (s.corr == -1) {
    const int16_t temp = Q[127];
    for (int i = 127; i > 0; i--) Q[i] = Q[i - 1];
    I[0] = (int16_t)s.saved;
    s.saved = temp;
  }
}

/* 128-point forward FFT, radix-2 decimation in frequency, the fixed operation network documented in DESIGN.md (section 11), evaluated in
 * place WITHOUT the final reordering: element e (buf[(2e)*stride], buf[(2e+1)*stride]) ends up holding X[brev7(e)].
 * tw = (cos, -sin)(2 pi k/128). */
/* one stage s (0..6) of the network, butterflies t = t0, t0 + tstep, ... < 64 (the kernel splits t over its warps) */
AUX_HD void pp_fft128_stage(float *buf, int stride, const float *tw, int s, int t0, int tstep) {
  const int half = 64 >> s;
  AUX_UNROLLN(4) for (int t = t0; t < 64; t += tstep) {
    const int j = t & (half - 1);
    const int i0 = ((t - j) << 1) + j, i1 = i0 + half;
    float *a = buf + (2 * i0) * stride, *b = buf + (2 * i1) * stride;
    const float ar = a[0], ai = a[stride], br = b[0], bi = b[stride];
    const float wr = tw[2 * (j << s)], wi = tw[2 * (j << s) + 1];
    const float tr = ar - br, ti = ai - bi;
    a[0] = ar + br;
    a[stride] = ai + bi;
    const float p0 = tr * wr, p1 = ti * wi, p2 = tr * wi, p3 = ti * wr;
    b[0] = p0 - p1;
    b[stride] = p2 + p3;
  }
}
AUX_HD void pp_fft128(float *buf, int stride, const float *tw) {
  for (int s = 0; s < 7; s++) pp_fft128_stage(buf, stride, tw, s, 0, 1);
}

/* Power spectrum, scan and the detector's state machine, PP.cpp:89-118.  buf as left by pp_fft128. */
/* |element e|^2 = re^2 + im^2, two products and one sum (arm_cmplx_mag_squared_f32, PP.cpp:89) */
AUX_HD float pp_power(const float *buf, int stride, int e) {
  const float re = buf[(2 * e) * stride], im = buf[(2 * e + 1) * stride];
  const float p = re * re, q = im * im;
  return p + q;
}

/* Scan and the detector's state machine, PP.cpp:91-118.  pw[e*stride] = power of element e (bin brev7(e)). */
AUX_HD void pp_decide(const float *pw, int stride, PpState &s) {
  float average_power = 0.0f, maximum_power = 0.0f;
  int maxLine = 0;
  for (int i = 5; i < 123; i++) {
    const float v = pw[aux_brev7(i) * stride];
    average_power = average_power + v;
    if (v > maximum_power) { maxLine = i; maximum_power = v; }
  }
  average_power = average_power / 118.0f;
  if ((double)maximum_power > 10.0 * (double)average_power) { /* the product is exact in double: an exact compare */
    const float imbalance_ratio = maximum_power / pw[aux_brev7(128 - maxLine) * stride]; /* maxLine >= 5 here */
    if (imbalance_ratio < 10.0f) s.fail++;
    else s.fail = 0;
    if (s.fail > 10) {
      s.corr++;
      if (s.corr > 1) s.corr = -1;
      s.fail = 0;
      s.succ = 0;
    }
    s.succ++;
  }
  if (s.succ > 1000) s.autod = 0;
}

/* Power spectrum in place (slot e is read, as 2e and 2e+1 >= e, before it is written) + decisions: the one-thread form */
AUX_HD void pp_detect(float *buf, int stride, PpState &s) {
  for (int e = 0; e < 128; e++) buf[e * stride] = pp_power(buf, stride, e);
  pp_decide(buf, stride, s);
}

/* ------------------------------------------------------------------ I/Q generator ---- */
/* The kernel keeps a span of one channel's input as floats in shared memory at padded addresses: sample position q lives
 * at q + (q >> 4), so that 32 lanes reading positions 16 apart hit 32 different banks (stride 17). */
#define IQ_LANE_TILE 16 /* outputs per lane */
AUX_HD int iq_pad(int q) { return q + (q >> 4); }

/* acc[c] = sum_{k=0..63} h[k] * (x[n0+c-(2k+1)] - x[n0+c-255+2k]), k ascending, product and sum rounded separately
 * (IQ.cpp:65-73), for the 16 consecutive outputs n0..n0+15 of one lane; x[n0] sits at position q0 (q0 % 16 == 0,
 * q0 >= 256).  Two 16-register windows slide by two positions per tap pair; after 8 taps a window is entirely new, so
 * the 8-tap body is the loop body and all addresses are immediates off two pointers that move by 17 words per body. */
AUX_HD void iq_lane_fir(const float *xs, int q0, const float *h, float *acc) {
  const float *p = xs + iq_pad(q0);
  float a[16], b[16];
  a[0] = p[-2]; /* position q0-1: previous group */
  AUX_UNROLL for (int c = 1; c < 16; c++) a[c] = p[c - 1];
  AUX_UNROLL for (int c = 0; c < 15; c++) b[c] = p[c - 255 - 16];
  b[15] = p[-240 - 15];
  AUX_UNROLL for (int c = 0; c < 16; c++) acc[c] = 0.0f;
  const float *pa = p, *pb = p;
  AUX_UNROLLN(1) for (int m = 0; m < 8; m++) {
    AUX_UNROLL for (int u = 0; u < 8; u++) {
      const float hk = h[8 * m + u];
      AUX_UNROLL for (int c = 0; c < 16; c++) {
        const float d = a[c] - b[c];
        const float pr = hk * d;
        acc[c] = acc[c] + pr;
      }
      AUX_UNROLL for (int c = 15; c >= 2; c--) a[c] = a[c - 2];
      { /* positions q0 - 3 - 2k and q0 - 2 - 2k, k = 8m + u */
        const int s0 = -3 - 2 * u, s1 = -2 - 2 * u;
        a[0] = pa[s0 + (s0 >= -16 ? -1 : -2)];
        a[1] = pa[s1 - 1];
      }
      AUX_UNROLL for (int c = 0; c < 14; c++) b[c] = b[c + 2];
      { /* positions q0 - 239 + 2k and q0 - 238 + 2k */
        const int s0 = -239 + 2 * u, s1 = -238 + 2 * u;
        b[14] = pb[s0 - 15];
        b[15] = pb[s1 + (s1 >= -224 ? -14 : -15)];
      }
    }
    pa -= 17;
    pb += 17;
  }
}

/* the delayed I rail of the same 16 outputs: x[n0 + c - 128] (IQ.cpp:75) */
AUX_HD float iq_lane_delayed(const float *xs, int q0, int c) { return xs[iq_pad(q0) + c - 136]; }

#endif
