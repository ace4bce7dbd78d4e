/* oracle/sdr_oracle.c -- TEST INFRASTRUCTURE, not product code.  See sdr_oracle.h.
 *
 * CPU restatement of the reference chain, written to follow the reference's
 * arithmetic operation by operation (same operand order, same float/double
 * promotions, same truncations) so that it can be pinned bit-for-bit to the
 * host-compiled reference (oracle/_ref).  Build: plain -O2, FP contraction off.
 *
 * C: = /root/reference/SRC/AudioSDRlib/AudioSDR.cpp, H: = .../AudioSDR.h
 */
#define _GNU_SOURCE
#include "sdr_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sdr_oracle_tables.inc"

#define NB ORA_NBLOCK
#define ORA_PI 3.1415926535897932384626433832795 /* Arduino's PI, a double (shim Arduino.h) */
static const float FS = 44100.0f;                /* AUDIO_SAMPLE_RATE_EXACT on Teensy 4 */

enum { LSB = 0, USB = 1, CW_LSB = 2, CW_USB = 3, AM = 4, SAM = 5, WSPR = 6 };          /* H:44-50 */
enum { F_AM = 0, F_CW, F_WSPR, F_2100, F_2300, F_2500, F_2700, F_2900, F_3100, F_3300, F_BYPASS }; /* H:56-66 */

static inline float tabf(const uint32_t *t, int i) { float f; memcpy(&f, &t[i], 4); return f; }

typedef struct { const uint32_t *coef; float st[16]; } cascade; /* arm_biquad_casd_df1_inst_f32 + its state */

struct ora_channel {
  /* general (H:164-182) */
  float audio[NB], I[NB], Q[NB];
  float in_gain, in_gain_i, in_gain_q, gain_balance, out_gain;
  float freq_shift;
  int mode;
  int muted;
  /* biquads (H:184-195) */
  cascade if_i, if_q, img_i, img_q, aud;
  int aud_enabled, aud_id;
  /* function-statics of update() (C:41-44) */
  float hist_i[4 * NB], hist_q[4 * NB];
  float phase_ssb, phase_am;
  /* ALS (H:198-205) */
  int als_m, als_delay;
  float als_lambda, als_in[2 * NB], als_c[NB];
  int als_on, als_notch, als_adaptive;
  /* AGC (H:208-232) */
  float agc_carrier, a_att, a_rel, t_att, b_att, b_rel, agc_gain, lut[130], t_hang, knee, static_gain, slope, t_rel,
      thr, absval, old_absval;
  uint32_t hang_count, hang_ctr;
  int agc_active, agc_on;
  /* noise blanker (H:235-246) */
  float nb_i[3 * NB], nb_q[3 * NB], nb_mask[3 * NB];
  float nb_alpha, nb_beta, nb_thr, nb_mag, nb_avg;
  int nb_pre, nb_post, nb_on, nb_hit;
  /* SAM (H:249-284 members, C:690-694 statics) */
  float s_alpha, s_beta, s_fconv, lock_lo, lock_hi, b0, b1, a1;
  float y_re, y_im, prev_filt, d0, d1, phase_est, pll_f;
  int locked;
  float *taps;
};

/* ---- CMSIS-DSP arm_biquad_cascade_df1_{init_,}f32 restated (see oracle/shim/arm_math.h header note) ---- */
static void cascade_init(cascade *c, const uint32_t *coef) { c->coef = coef; memset(c->st, 0, sizeof c->st); }

static void cascade_run(cascade *c, const float *src, float *dst) {
  const float *in = src;
  for (int s = 0; s < 4; s++) {
    float b0 = tabf(c->coef, 5 * s), b1 = tabf(c->coef, 5 * s + 1), b2 = tabf(c->coef, 5 * s + 2),
          a1 = tabf(c->coef, 5 * s + 3), a2 = tabf(c->coef, 5 * s + 4);
    float x1 = c->st[4 * s], x2 = c->st[4 * s + 1], y1 = c->st[4 * s + 2], y2 = c->st[4 * s + 3];
    for (int n = 0; n < NB; n++) {
      float x = in[n];
      float acc = (b0 * x) + (b1 * x1) + (b2 * x2) + (a1 * y1) + (a2 * y2);
      dst[n] = acc;
      x2 = x1; x1 = x; y2 = y1; y1 = acc;
    }
    c->st[4 * s] = x1; c->st[4 * s + 1] = x2; c->st[4 * s + 2] = y1; c->st[4 * s + 3] = y2;
    in = dst;
  }
}

/* ---- sine LUT oscillator (H:358-377) ---- */
static float lut_sin(float ph) {
  const float two_pi = (float)(2.0 * ORA_PI);
  if (ph >= two_pi) ph -= two_pi;
  if ((double)ph < 0.0) ph += two_pi;
  uint16_t ip = (uint16_t)(long)((double)ph * 65535.0 / (double)two_pi);
  uint16_t idx = ip >> 8, frac = ip & 0xFF;
  float v1 = tabf(SDR_TAB_SINE, idx), v2 = tabf(SDR_TAB_SINE, idx + 1);
  float prod = (v2 - v1) * (float)frac;
  return (float)((double)v1 + (double)prod / 256.0);
}
static float lut_cos(float ph) { return lut_sin((float)((double)ph + ORA_PI / 2.0)); }

/* ---- complex frequency shifter (H:508-526) ---- */
static float shifter(float *I, float *Q, float f_shift, float phase) {
  const float two_pi = (float)(2.0 * ORA_PI);
  float inc = f_shift * (two_pi / FS);
  for (int n = 0; n < NB; n++) {
    float c = lut_cos(phase), s = lut_sin(phase);
    float ti = I[n], tq = Q[n];
    I[n] = ti * c - tq * s;
    Q[n] = tq * c + ti * s;
    phase += inc;
    if (phase > two_pi) phase -= two_pi;
    else if ((double)phase < 0.0) phase += two_pi;
  }
  return phase;
}

/* ---- atan2 approximation (H:384-408) ---- */
static float atan_poly(float z) { return (0.97239411f + -0.19194795f * z * z) * z; }
static float atan2_approx(float y, float x) {
  const float half_pi = (float)(0.5 * ORA_PI);
  if ((double)x != 0.0) {
    if (fabsf(x) > fabsf(y)) {
      float z = y / x;
      if ((double)x > 0.0) return atan_poly(z);
      else if ((double)y >= 0.0) return (float)((double)atan_poly(z) + ORA_PI);
      else return (float)((double)atan_poly(z) - ORA_PI);
    } else {
      float z = x / y;
      if ((double)y > 0.0) return -atan_poly(z) + half_pi;
      else return -atan_poly(z) - half_pi;
    }
  } else {
    if ((double)y > 0.0) return half_pi;
    else if ((double)y < 0.0) return -half_pi;
  }
  return 0.0f;
}

/* ---- bit-hack square root with one Newton step (H:434-446, called with n_iter = 1 at C:628) ---- */
static float sqrt_hack(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u -= 1u << 23; u >>= 1; u += 1u << 29;
  float out; memcpy(&out, &u, 4);
  out = (float)(0.5 * (double)(out + x / out));
  return out;
}

/* ---- log2 approximation (H:483-491), used only by the AGC table builder ---- */
static float log2_approx(float v) {
  int e; float m = frexpf(fabsf(v), &e);
  return (((1.23149591368684f * m - 4.11852516267426f) * m + 6.02197014179219f) * m - 3.13396450166353f) + e;
}

/* ---- AGC: table builder (C:459-480), time constants (C:551-566), init (C:439-457) ---- */
static void agc_build_lut(ora_channel *c) {
  float lo = expf(2.3025 * (c->thr - c->knee / 2.0) / 20.0);
  float hi = expf(2.3025 * (c->thr + c->knee / 2.0) / 20.0);
  for (int i = 0; i < 130; i++) { /* the reference writes 130 entries into a 129-float array (Q8); entry 129 is never read */
    float in = (float)i / 128.0;
    float in_db = 6.026 * log2_approx(in);
    if (in < lo) c->lut[i] = 1.0;
    else if (in > hi) {
      float out_db = (c->thr + (in_db - c->thr) * c->slope);
      c->lut[i] = expf(2.3025 * (out_db - in_db) / 20.0);
    } else {
      float out_db = in_db + ((c->slope - 1.0) * (in_db - c->thr + c->knee / 2.0) * (in_db - c->thr + c->knee / 2.0)) /
                                 (2.0 * c->knee);
      c->lut[i] = expf(2.3025 * (out_db - in_db) / 20.0);
    }
  }
}
static void agc_set_attack(ora_channel *c, float ms) {
  c->t_att = ms; c->a_att = exp(log(0.1) / (FS * c->t_att / 1000.0)); c->b_att = 1.0 - c->a_att;
}
static void agc_set_release(ora_channel *c, float ms) {
  c->t_rel = ms; c->a_rel = exp(log(0.1) / (FS * c->t_rel / 1000.0)); c->b_rel = 1.0 - c->a_rel;
}
static void agc_set_hang(ora_channel *c, float ms) { c->t_hang = ms; c->hang_count = c->t_hang * FS / 1000.0; }
static void agc_init(ora_channel *c) {
  c->thr = -60.0; c->slope = 0.1; c->knee = 2.0; c->t_att = 5.0; c->t_rel = 500.0; c->t_hang = 100.0;
  c->hang_count = FS * (c->t_hang / 1000.0);
  c->a_att = exp(log(0.1) / (FS * c->t_att / 1000.0)); c->b_att = 1.0 - c->a_att;
  c->a_rel = exp(log(0.1) / (FS * c->t_rel / 1000.0)); c->b_rel = 1.0 - c->a_rel;
  c->agc_on = 1;
  agc_build_lut(c);
  c->t_hang = c->lut[129]; /* Q8: the 130th table write lands on _agc_hangTime; no effect on the audio */
}
static float agc_lookup(const ora_channel *c, uint16_t v) { /* C:483-494 */
  uint16_t idx = v >> 8; if (idx > 127) idx = 127;
  uint16_t frac = v & 0xFF;
  float d = (float)frac / 256.0;
  return c->lut[idx] + (c->lut[idx + 1] - c->lut[idx]) * d;
}
static void agc_run(ora_channel *c, float *b) { /* C:404-436 */
  for (int n = 0; n < NB; n++) {
    if (c->mode == AM) c->absval = 2.0 * c->agc_carrier;
    else c->absval = fabsf(b[n]);
    if ((double)c->absval > 1.0) c->absval = 1.0;
    if (c->absval > c->old_absval) {
      c->absval = c->a_att * c->old_absval + c->b_att * c->absval;
      c->old_absval = c->absval;
      c->hang_ctr = c->hang_count;
      c->agc_gain = agc_lookup(c, (uint16_t)(int)((double)c->absval * 32767.0));
    } else {
      if (c->hang_ctr > 0) c->hang_ctr--;
      else {
        c->absval = c->a_rel * c->old_absval + c->b_rel * c->absval;
        c->old_absval = c->absval;
        c->agc_gain = agc_lookup(c, (uint16_t)(int)((double)c->absval * 32767.0));
      }
    }
    c->agc_active = ((double)c->agc_gain < 0.99);
    float o = c->agc_gain * c->static_gain * b[n];
    o = ((double)o > 1.0) ? 1.0f : o;
    o = ((double)o < -1.0) ? -1.0f : o;
    b[n] = o;
  }
}

/* ---- impulse noise blanker (C:606-650), rings reset by C:676-682 ---- */
static void nb_reset(ora_channel *c) {
  for (int i = 0; i < 3 * NB; i++) { c->nb_i[i] = 0.0f; c->nb_q[i] = 0.0f; c->nb_mask[i] = 1.0f; }
}
static void nb_run(ora_channel *c, float *I, float *Q) {
  static const float dn[7] = {0.933, 0.750, 0.500, 0.250, 0.067, 0.000, 0.000};
  static const float up[7] = {0.000, 0.000, 0.067, 0.250, 0.500, 0.750, 0.933};
  c->nb_hit = 0;
  for (int i = 0; i < NB; i++) {
    c->nb_i[i] = c->nb_i[NB + i]; c->nb_i[NB + i] = c->nb_i[2 * NB + i]; c->nb_i[2 * NB + i] = I[i];
    c->nb_q[i] = c->nb_q[NB + i]; c->nb_q[NB + i] = c->nb_q[2 * NB + i]; c->nb_q[2 * NB + i] = Q[i];
    c->nb_mask[i] = c->nb_mask[NB + i]; c->nb_mask[NB + i] = c->nb_mask[2 * NB + i]; c->nb_mask[2 * NB + i] = 1.0f;
  }
  for (int i = NB - 50; i < 2 * NB; i++) {
    c->nb_mag = sqrt_hack(c->nb_i[i] * c->nb_i[i] + c->nb_q[i] * c->nb_q[i]);
    if (c->nb_mag > c->nb_avg * c->nb_thr) {
      for (int j = -c->nb_pre; j < c->nb_post + 1; j++) c->nb_mask[i + j] = 0.0f;
      c->nb_hit = 1;
    }
    c->nb_avg = c->nb_alpha * c->nb_avg + c->nb_beta * c->nb_mag;
  }
  for (int i = NB; i < 2 * NB; i++) {
    if ((c->nb_mask[i] == 1.0f) && (c->nb_mask[i - 1] == 0.0f)) {
      for (int j = 0; j < 7; j++) c->nb_mask[i - 7 + j] = dn[j];
    } else if ((c->nb_mask[i - 1] == 0.0f) && (c->nb_mask[i] == 1.0f)) { /* same condition: dead branch (Q3) */
      for (int j = 0; j < 7; j++) c->nb_mask[i + j - 1] = up[j];
    }
  }
  for (int i = 0; i < NB; i++) { I[i] = c->nb_mask[i] * c->nb_i[i]; Q[i] = c->nb_mask[i] * c->nb_q[i]; }
}

/* ---- ALS LMS notch / peak filter (C:324-352) ---- */
static void als_run(ora_channel *c, float *b) {
  uint16_t count = 0;
  for (int i = 0; i < NB; i++) { c->als_in[i] = c->als_in[NB + i]; c->als_in[NB + i] = b[i]; }
  for (int i = NB; i < 2 * NB; i++) {
    float y = 0.0f;
    for (int j = 0; j < c->als_m; j++) y += c->als_c[j] * c->als_in[(i - c->als_delay) - j];
    float e = c->als_in[i] - y;
    if (c->als_adaptive) {
      if (count == 0)
        for (int j = 0; j < c->als_m; j++) {
          float g = e * c->als_in[i - c->als_delay - j];
          c->als_c[j] += c->als_lambda * g;
        }
      count = (count + 1) % 4;
    }
    b[i - NB] = c->als_notch ? e : y;
  }
}

/* ---- synchronous AM PLL (C:688-749) ---- */
static void sam_run(ora_channel *c, float *I, float *Q) {
  const float two_pi = (float)(2.0 * ORA_PI);
  for (int n = 0; n < NB; n++) {
    float xr = I[n], xi = Q[n];
    float dr = xr * c->y_re + xi * c->y_im;
    float di = xi * c->y_re - xr * c->y_im;
    float err = atan2_approx(di, dr);
    c->d1 = c->d0;
    c->d0 = err - c->a1 * c->d1;
    float filt = c->b0 * c->d0 + c->b1 * c->d1;
    c->phase_est = (float)((double)c->phase_est + (double)(filt + c->prev_filt) / 2.0);
    c->prev_filt = filt;
    while ((double)c->phase_est >= ORA_PI) c->phase_est -= two_pi;
    while ((double)c->phase_est < -ORA_PI) c->phase_est += two_pi;
    c->y_re = lut_cos(c->phase_est);
    c->y_im = lut_sin(c->phase_est);
    c->pll_f = c->s_alpha * c->pll_f + c->s_beta * (filt * c->s_fconv);
    c->locked = (c->pll_f > c->lock_lo) && (c->pll_f < c->lock_hi);
    if (c->locked) {
      float ti = I[n], tq = Q[n];
      I[n] = ti * c->y_re + tq * c->y_im;
      Q[n] = -ti * c->y_im + tq * c->y_re;
    }
  }
}

/* ---- setDemodMode (C:187-222) ---- */
static void set_mode(ora_channel *c, int m) {
  c->mode = (uint16_t)m;
  const float fc = 6890.0f, bw_ssb = 3000.0f, bw_cw = 1000.0f;
  const uint32_t *t = 0;
  if (c->mode == USB) { c->freq_shift = fc - bw_ssb / 2.0; t = SDR_TAB_IF_SSB; }
  else if (c->mode == LSB) { c->freq_shift = fc + bw_ssb / 2.0; t = SDR_TAB_IF_SSB; }
  else if (c->mode == WSPR) { c->freq_shift = fc - bw_ssb / 2.0; t = SDR_TAB_IF_WSPR; }
  else if (c->mode == CW_USB) { c->freq_shift = fc - bw_cw / 2.0; t = SDR_TAB_IF_CW; }
  else if (c->mode == CW_LSB) { c->freq_shift = fc + bw_cw / 2.0; t = SDR_TAB_IF_CW; }
  else if (c->mode == AM || c->mode == SAM) { c->freq_shift = fc; t = SDR_TAB_IF_AM; }
  if (t) { cascade_init(&c->if_i, t); cascade_init(&c->if_q, t); } /* IF state only; nothing else is reset */
}

static const uint32_t *audio_table(int id) {
  switch (id) {
    case F_AM: return SDR_TAB_AUDIO_AM; case F_CW: return SDR_TAB_AUDIO_CW; case F_WSPR: return SDR_TAB_AUDIO_WSPR;
    case F_2100: return SDR_TAB_AUDIO_2100; case F_2300: return SDR_TAB_AUDIO_2300; case F_2500: return SDR_TAB_AUDIO_2500;
    case F_2700: return SDR_TAB_AUDIO_2700; case F_2900: return SDR_TAB_AUDIO_2900; case F_3100: return SDR_TAB_AUDIO_3100;
    case F_3300: return SDR_TAB_AUDIO_3300; default: return 0;
  }
}

/* ---- init() (C:174-185) on top of the in-class member initialisers (H:164-284) ---- */
static void init(ora_channel *c) {
  cascade_init(&c->aud, SDR_TAB_AUDIO_2700);
  cascade_init(&c->if_i, SDR_TAB_IF_SSB);
  cascade_init(&c->if_q, SDR_TAB_IF_SSB);
  cascade_init(&c->img_i, SDR_TAB_AM_IMAGE);
  cascade_init(&c->img_q, SDR_TAB_AM_IMAGE);
  agc_init(c);
  nb_reset(c);
  set_mode(c, LSB);
  c->muted = 0;
}

ora_channel *ora_new(void) {
  ora_channel *c = (ora_channel *)calloc(1, sizeof *c); /* Q4: never-initialised members read as zero */
  c->in_gain = c->in_gain_i = c->in_gain_q = c->gain_balance = 1.0f; c->out_gain = 0.5; c->muted = 1;
  c->als_m = 55; c->als_delay = 3; c->als_lambda = 0.5; c->als_on = 0; c->als_notch = 1; c->als_adaptive = 1;
  c->t_hang = 100.0; c->static_gain = 10.0; c->agc_active = 1; c->agc_on = 1;
  c->nb_alpha = 0.995; c->nb_beta = (1.0 - c->nb_alpha); c->nb_thr = 1.2; c->nb_mag = 0.0; c->nb_avg = 10.0;
  c->nb_pre = 10; c->nb_post = 10; c->nb_on = 1; c->nb_hit = 0;
  {
    const float two_pi = 2.0 * ORA_PI; /* H:249 */
    c->s_alpha = 0.995; c->s_beta = 1.0 - c->s_alpha; c->s_fconv = (FS / two_pi);
    c->lock_lo = 6890.0f - 1000.0; c->lock_hi = 6890.0f + 1000.0;
    float wn = 0.07f, zeta = 0.707f, Ka = 1000.f;
    float tau1 = Ka / (wn * wn), tau2 = 2 * zeta / wn;
    c->b0 = (2 * Ka / tau1) * (1.0 + 2.0 * tau2);
    c->b1 = (2 * Ka / tau1) * (1.0 - 2.0 * tau2);
    c->a1 = -1.0;
  }
  init(c);
  return c;
}
void ora_free(ora_channel *c) { free(c); }
void ora_set_taps(ora_channel *c, float *taps) { c->taps = taps; }
void ora_get_phases(const ora_channel *c, float *ps, float *pa) { *ps = c->phase_ssb; *pa = c->phase_am; }

/* ---- the setter surface (C:187-311, 356-398, 498-566, 653-682) ---- */
int ora_apply(ora_channel *c, uint32_t op, float a0, float a1, float a2) {
  switch (op) {
    case 1: c->muted = (a0 != 0.0f); break;
    case 2: { float g = a0; if ((double)g > 10.0) g = 10.0; if ((double)g < 0.0) g = 0.0;         /* C:232-238 */
      c->in_gain = g; c->in_gain_i = c->in_gain * c->gain_balance; c->in_gain_q = c->in_gain; } break;
    case 3: { float bal = sqrtf(a0);                                                              /* C:240-244; the member stays 1 (Q7) */
      c->in_gain_i = c->in_gain * bal; c->in_gain_q = c->in_gain / bal; } break;
    case 4: set_mode(c, (int)a0); break;
    case 5: c->aud_enabled = 1; break;
    case 6: c->aud_enabled = 0; break;
    case 7: c->out_gain = a0; break;
    case 8: { int id = (int)a0;                                                                   /* C:298-311 */
      if (id == F_BYPASS) c->aud_enabled = 0;
      else { const uint32_t *t = audio_table(id); if (t) cascade_init(&c->aud, t); }
      c->aud_id = (int16_t)id; } break;
    case 9: c->als_on = 1; memset(c->als_c, 0, sizeof c->als_c); memset(c->als_in, 0, sizeof c->als_in); break; /* C:384-391 */
    case 10: c->als_on = 0; break;
    case 11: c->als_notch = 1; break;
    case 12: c->als_notch = 0; break;
    case 13: c->als_adaptive = 1; break;
    case 14: c->als_adaptive = 0; break;
    case 15: c->als_m = (int16_t)(unsigned)a0; if (c->als_m >= NB) c->als_m = NB;                 /* C:393-398 */
      c->als_lambda = a1; c->als_delay = (int16_t)a2; break;
    case 16: c->agc_on = 1; break;
    case 17: c->agc_on = 0; break;
    case 18: c->thr = a0; agc_build_lut(c); c->t_hang = c->lut[129]; break;
    case 19: c->slope = a0; agc_build_lut(c); c->t_hang = c->lut[129]; break;
    case 20: { int m = (int16_t)a0;                                                               /* C:524-544 */
      if (m == 0) c->agc_on = 0;
      else if (m == 1) { agc_set_attack(c, 2.0); agc_set_release(c, 100.0); agc_set_hang(c, 100.0); c->agc_on = 1; }
      else if (m == 2) { agc_set_attack(c, 5.0); agc_set_release(c, 250.0); agc_set_hang(c, 500.0); c->agc_on = 1; }
      else if (m == 3) { agc_set_attack(c, 10.0); agc_set_release(c, 500.0); agc_set_hang(c, 2000.0); c->agc_on = 1; } } break;
    case 21: c->knee = a0; agc_build_lut(c); c->t_hang = c->lut[129]; break;
    case 22: agc_set_attack(c, a0); break;
    case 23: agc_set_release(c, a0); break;
    case 24: agc_set_hang(c, a0); break;
    case 25: c->static_gain = a0; break;
    case 26: c->nb_on = 1; nb_reset(c); break;
    case 27: c->nb_on = 0; break;
    case 28: c->nb_thr = a0; nb_reset(c); break;
    case 29: c->nb_thr = powf(10.0, (a0 / 20.0)); nb_reset(c); break;
    case 30: init(c); break;
    case 100: { static const uint32_t ident[20] = {0x3F800000u, 0, 0, 0, 0, 0x3F800000u, 0, 0, 0, 0,
                                                    0x3F800000u, 0, 0, 0, 0, 0x3F800000u, 0, 0, 0, 0};
      c->if_i.coef = ident; c->if_q.coef = ident; } break;
    default: return -1;
  }
  return 0;
}

static void tap(ora_channel *c, int id, const float *v) { if (c->taps) memcpy(c->taps + id * NB, v, NB * sizeof(float)); }

/* ---- update() body after input scaling (C:71-168) ---- */
static void chain(ora_channel *c, float *audio_out, int16_t *pcm) {
  float *I = c->I, *Q = c->Q, *A = c->audio;
  tap(c, ORA_TAP_IN_I, I); tap(c, ORA_TAP_IN_Q, Q);
  if (c->nb_on) nb_run(c, I, Q);                                                                   /* C:73 */
  tap(c, ORA_TAP_NB_I, I); tap(c, ORA_TAP_NB_Q, Q);
  cascade_run(&c->if_i, I, I);                                                                     /* C:77-78 */
  cascade_run(&c->if_q, Q, Q);
  tap(c, ORA_TAP_IF_I, I); tap(c, ORA_TAP_IF_Q, Q);
  int m = c->mode;
  if (m == USB || m == LSB || m == CW_USB || m == CW_LSB || m == WSPR) {                           /* C:84-119 */
    c->phase_ssb = shifter(I, Q, -c->freq_shift, c->phase_ssb);
    for (int i = 0; i < NB; i++) {
      c->hist_i[i] = c->hist_i[NB + i]; c->hist_i[NB + i] = c->hist_i[2 * NB + i];
      c->hist_i[2 * NB + i] = c->hist_i[3 * NB + i]; c->hist_i[3 * NB + i] = I[i];
      c->hist_q[i] = c->hist_q[NB + i]; c->hist_q[NB + i] = c->hist_q[2 * NB + i];
      c->hist_q[2 * NB + i] = c->hist_q[3 * NB + i]; c->hist_q[3 * NB + i] = Q[i];
    }
    for (int i = 0; i < NB; i++) {
      Q[i] = 0.0f;
      for (int k = 0; k < 64; k++) {
        int i1 = (3 * NB + i) - (2 * k + 1);
        int i2 = (3 * NB + i) - 257 + 2 * (k + 1);
        Q[i] += tabf(SDR_TAB_HILBERT, k) * (c->hist_q[i1] - c->hist_q[i2]);
      }
      I[i] = c->hist_i[3 * NB + i - 128];
    }
    for (int i = 0; i < NB; i++) {
      if (m == USB || m == CW_USB || m == WSPR) A[i] = I[i] - Q[i];
      else A[i] = I[i] + Q[i];
    }
  } else if (m == AM || m == SAM) {                                                                /* C:122-144 */
    if (m == SAM) { sam_run(c, I, Q); for (int i = 0; i < NB; i++) A[i] = Q[i]; }
    if (m == AM || (m == SAM && !c->locked)) {
      c->phase_am = shifter(I, Q, -6890.0f, c->phase_am);
      cascade_run(&c->img_i, I, I);
      cascade_run(&c->img_q, Q, Q);
      for (int i = 0; i < NB; i++) {
        A[i] = sqrtf(I[i] * I[i] + Q[i] * Q[i]);
        float aa = (A[i] > 0) ? A[i] : -A[i];
        c->agc_carrier = (float)(.995 * (double)c->agc_carrier + 0.005 * (double)aa);
      }
    }
  }
  tap(c, ORA_TAP_DM_I, I); tap(c, ORA_TAP_DM_Q, Q); tap(c, ORA_TAP_DEMOD, A);
  if (c->aud_enabled) { float tmp[NB]; memcpy(tmp, A, sizeof tmp); cascade_run(&c->aud, tmp, A); } /* C:149, 280-286 */
  tap(c, ORA_TAP_AUDF, A);
  if (c->agc_on) agc_run(c, A);                                                                    /* C:152 */
  tap(c, ORA_TAP_AGC, A);
  if (c->als_on) als_run(c, A);                                                                    /* C:155 */
  tap(c, ORA_TAP_ALS, A);
  for (int i = 0; i < NB; i++) {                                                                   /* C:158-161 */
    float g = c->out_gain * A[i];
    if (audio_out) audio_out[i] = c->muted ? 0.0f : g;
    if (pcm) pcm[i] = c->muted ? 0 : (int16_t)(int)((double)g * 32767.0);
  }
}

void ora_update_i16(ora_channel *c, const int16_t *I, const int16_t *Q, float *audio, int16_t *pcm) {
  for (int i = 0; i < NB; i++) {                                                                   /* C:67-70 */
    c->I[i] = (float)(((double)(float)I[i] / 32767.0) * (double)c->in_gain_i);
    c->Q[i] = (float)(((double)(float)Q[i] / 32767.0) * (double)c->in_gain_q);
  }
  chain(c, audio, pcm);
}

void ora_update_f32(ora_channel *c, const float *I, const float *Q, float *audio, int16_t *pcm) {
  for (int i = 0; i < NB; i++) { /* float32 boundary: x stands for q/32767; equals C:67-70 whenever the gain is 1 */
    c->I[i] = (float)((double)I[i] * (double)c->in_gain_i);
    c->Q[i] = (float)((double)Q[i] * (double)c->in_gain_q);
  }
  chain(c, audio, pcm);
}

void ora_status(const ora_channel *c, float *s) {
  memset(s, 0, sizeof(float) * ORA_NSTATUS);
  const float fc = 6890.0f;
  int m = c->mode;
  s[0] = c->freq_shift; s[1] = (float)m; s[2] = (float)c->agc_active; s[3] = (float)c->nb_hit;
  s[4] = c->pll_f; s[5] = (float)c->locked; s[6] = c->agc_carrier;
  /* getBPFlower / getBPFupper (C:259-273; WSPR upper has the `+-` typo, Q11) */
  if (m == USB || m == LSB) { s[7] = fc - 3000.0f / 2.0; s[8] = fc + 3000.0f / 2.0; }
  else if (m == CW_USB || m == CW_LSB) { s[7] = fc - 1000.0f / 2.0; s[8] = fc + 1000.0f / 2.0; }
  else if (m == AM || m == SAM) { s[7] = fc - 8500.0f / 2.0; s[8] = fc + 8500.0f / 2.0; }
  else if (m == WSPR) { s[7] = fc - 1000.0f / 2.0; s[8] = fc + -1000.0f / 2.0; }
  s[9] = (float)c->muted; s[10] = (float)c->aud_id; s[11] = (float)c->agc_on; s[12] = (float)c->nb_on;
  s[13] = (float)c->als_on; s[14] = c->agc_gain; s[15] = c->nb_avg;
}

/* ---- batch helper ---- */
typedef struct {
  uint32_t c0, c1, n_blocks, n_events; const ora_event *ev; int fmt; const void *I, *Q;
  float *audio; int16_t *pcm; float *status; double secs;
} job;

static int ev_cmp(const void *a, const void *b) {
  const ora_event *x = *(const ora_event *const *)a, *y = *(const ora_event *const *)b;
  if (x->block != y->block) return x->block < y->block ? -1 : 1;
  return x < y ? -1 : (x > y ? 1 : 0); /* stable: file order within a block */
}

static void *worker(void *arg) {
  job *j = (job *)arg;
  size_t ns = (size_t)j->n_blocks * NB;
  const ora_event **mine = (const ora_event **)malloc(sizeof(void *) * (j->n_events + 1));
  double secs = 0;
  for (uint32_t ch = j->c0; ch < j->c1; ch++) {
    uint32_t ne = 0;
    for (uint32_t e = 0; e < j->n_events; e++)
      if (j->ev[e].channel == ch || j->ev[e].channel == 0xFFFFFFFFu) mine[ne++] = &j->ev[e];
    qsort(mine, ne, sizeof(void *), ev_cmp);
    ora_channel *c = ora_new();
    uint32_t ei = 0;
    struct timespec t0, t1;
    for (uint32_t b = 0; b < j->n_blocks; b++) {
      while (ei < ne && mine[ei]->block <= b) { ora_apply(c, mine[ei]->opcode, mine[ei]->a0, mine[ei]->a1, mine[ei]->a2); ei++; }
      float *ao = j->audio ? j->audio + ch * ns + (size_t)b * NB : 0;
      int16_t *po = j->pcm ? j->pcm + ch * ns + (size_t)b * NB : 0;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      if (j->fmt == 0)
        ora_update_i16(c, (const int16_t *)j->I + ch * ns + (size_t)b * NB, (const int16_t *)j->Q + ch * ns + (size_t)b * NB, ao, po);
      else
        ora_update_f32(c, (const float *)j->I + ch * ns + (size_t)b * NB, (const float *)j->Q + ch * ns + (size_t)b * NB, ao, po);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      secs += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    }
    while (ei < ne) { ora_apply(c, mine[ei]->opcode, mine[ei]->a0, mine[ei]->a1, mine[ei]->a2); ei++; }
    if (j->status) ora_status(c, j->status + (size_t)ch * ORA_NSTATUS);
    ora_free(c);
  }
  free(mine);
  j->secs = secs;
  return 0;
}

double ora_run(uint32_t n_channels, uint32_t n_blocks, const ora_event *events, uint32_t n_events, int fmt,
               const void *I, const void *Q, float *audio, int16_t *pcm, float *status, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((uint32_t)n_threads > n_channels) n_threads = (int)n_channels;
  if (n_threads < 1) return 0.0;
  job *jobs = (job *)calloc(n_threads, sizeof(job));
  pthread_t *th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
  for (int t = 0; t < n_threads; t++) {
    jobs[t] = (job){(uint32_t)((uint64_t)n_channels * t / n_threads), (uint32_t)((uint64_t)n_channels * (t + 1) / n_threads),
                    n_blocks, n_events, events, fmt, I, Q, audio, pcm, status, 0.0};
    if (n_threads == 1) worker(&jobs[t]);
    else pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  double mx = 0;
  for (int t = 0; t < n_threads; t++) {
    if (n_threads > 1) pthread_join(th[t], 0);
    if (jobs[t].secs > mx) mx = jobs[t].secs;
  }
  free(jobs); free(th);
  return mx;
}
