/* sdr_types.h -- PODs shared by the host side (sdr_host.cpp) and the kernels (sdr_kernel.cu).
 *
 * HBM layout
 * ----------
 * state   : float/uint32 words, CHANNEL-FASTEST:  word w of channel c is state[w * ch_stride + c].
 *           A warp whose 32 lanes own 32 consecutive channels therefore touches one 128-byte line per
 *           state word.  The state IS the checkpoint: everything a channel carries from block to block.
 * cfg     : one SdrChanCfg per channel (array of structs; read once per launch per role).
 * groups  : one SdrGroup per CTA: 32 channel ids of one pipeline class (-1 = empty lane).
 * planes  : caller-owned I/Q/audio, channel-major (include/sdr_batch.h).
 */
#ifndef SDR_TYPES_H
#define SDR_TYPES_H
#include <stdint.h>

#define SDR_T 32        /* samples per pipeline tile; 4 tiles = one reference block of 128 */
#define SDR_LANES 32    /* channels per group (one warp lane each) */
#define SDR_TPB 4       /* tiles per block */

/* ---- per-channel state words (reference member it stands for) ---- */
enum {
  W_IF_I = 0,        /* 16: _IFfilterStateI        H:191  {x1,x2,y1,y2} x 4 stages */
  W_IF_Q = 16,       /* 16: _IFfilterStateQ        H:192 */
  W_IMG_I = 32,      /* 16: _AMimage_stateI        H:193 */
  W_IMG_Q = 48,      /* 16: _AMimage_stateQ        H:194 */
  W_AUD = 64,        /* 16: _audio_filter_state    H:195 */
  W_PH_SSB = 80,     /* phase_SSB                  C:43 */
  W_PH_AM = 81,      /* phase_AM                   C:44 */
  W_AGC_GAIN = 82,   /* _agc_gain                  H:217 */
  W_AGC_OLD = 83,    /* _old_absVal                H:227 */
  W_AGC_HANG = 84,   /* _agc_hang_counter (u32)    H:229 */
  W_AGC_ACTIVE = 85, /* _agc_is_active (u32)       H:230 */
  W_AGC_CARRIER = 86,/* _agc_AMcarrierLevel        H:210 */
  W_SAM_YRE = 87,    /* yReal (function static)    C:690 */
  W_SAM_YIM = 88,    /* yImag                      C:691 */
  W_SAM_PREV = 89,   /* prev_phase_err_filt        C:692 */
  W_SAM_D0 = 90,     /* delay0                     H:270 */
  W_SAM_D1 = 91,     /* delay1 */
  W_SAM_PHASE = 92,  /* phase_est                  H:271 */
  W_SAM_FREQ = 93,   /* _PLLfreq                   H:273 */
  W_SAM_LOCKED = 94, /* _SAM_PLL_isLocked (u32)    H:274 */
  W_NB_AVG = 95,     /* _nb_AvgMag                 H:242 */
  W_NB_HIT = 96,     /* _nb_impulseDetected (u32)  H:246 */
  W_HQ = 128,        /* 256: the last 256 down-converted Q samples, oldest first (bufferQ, C:42) */
  W_HI = 384,        /* 128: the last 128 down-converted I samples (bufferI, C:41) */
  W_ALS_C = 512,     /* 128: _als_coeffs           H:202 */
  W_ALS_H = 640,     /* 128: the last 128 ALS inputs (_als_in[0..127] after the shift, H:201) */
  W_NB_MASK = 768,   /* 96 words = 384 byte codes: _mask, block slot (abs_block % 3), H:237 */
  W_NB_RING = 864,   /* 3 x 384: _BufferI, _BufferQ (H:235-236) and the envelope sqrt(I^2+Q^2) of every ring sample
                        (computed once on arrival instead of at each of its two scans), 3 block slots of 128, slot = abs_block % 3 */
  SDR_STATE_WORDS = 2016
};

/* noise-blanker mask codes (byte) -> value; code 0 must be 1.0 so that zeroed state == initBlanker() */
enum { MK_ONE = 0, MK_ZERO = 1, MK_933 = 2, MK_750 = 3, MK_500 = 4, MK_250 = 5, MK_067 = 6 };

/* ---- per-channel device configuration (resolved on the host from the setter shadow) ---- */
enum {
  CF_NB = 1u, CF_AUD = 2u, CF_AGC = 4u, CF_ALS = 8u, CF_ALS_NOTCH = 16u, CF_ALS_ADAPT = 32u, CF_MUTED = 64u,
  GF_LUT_GLOBAL = 0x10000u /* group summary only: some lane's AGC table is not among the 4 staged ones */
};
typedef struct {
  int32_t mode;            /* SDR_LSB..SDR_WSPR */
  uint32_t flags;          /* CF_* */
  float in_gain_i, in_gain_q, out_gain;
  float ssb_phase_inc;     /* (-_freq_shift) * (twoPI / fs), float ops as H:510 */
  int32_t if_set;          /* index into the IF coefficient sets {SSB, CW, WSPR, AM} */
  int32_t aud_set;         /* index into the audio sets (SDR_AUDIO_AM..SDR_AUDIO_3300) */
  float agc_a_att, agc_b_att, agc_a_rel, agc_b_rel, agc_static_gain;
  uint32_t agc_hang_count;
  int32_t agc_lut;         /* index into the handle's table of distinct 130-entry AGC tables */
  float nb_thr;
  int32_t als_m, als_delay;
  float als_lambda;
  uint32_t pad;
} SdrChanCfg;

/* ---- pipeline classes ---- */
enum { CLS_SSB = 0 /* LSB USB CW_LSB CW_USB WSPR: NCO + Hilbert */, CLS_ENV = 1 /* AM SAM: PLL + envelope */ };

#define SDR_LUT_SLOTS 4
typedef struct {
  int32_t cls;
  uint32_t feat;                 /* OR of the lanes' CF_* flags */
  int32_t cid[SDR_LANES];        /* channel id per lane, -1 = empty */
  int32_t lut_ids[SDR_LUT_SLOTS];/* up to 4 distinct AGC tables of this group, staged in shared memory (-1 = unused) */
  uint8_t lut_slot[SDR_LANES];   /* per lane: slot in lut_ids, or 255 = read the table from global memory */
} SdrGroup;

/* ---- constant tables in device memory ---- */
typedef struct {
  float if_sets[4][20];
  float aud_sets[10][20];
  float am_image[20];
  float hilbert[64];
  float sine[260];
  float pk_consts[8];      /* {1,1,-1,-1,-0,-0}: multiplicands / addends of the packed FMA forms, deliberately run-time data (sdr_pipeline.cuh, PkConst) */
} SdrTables;

typedef struct {
  const void *in_i, *in_q;
  void *out;
  unsigned long long in_pitch, out_pitch; /* elements */
  int32_t in_fmt, out_fmt;
  uint32_t n_tiles;      /* 4 * n_blocks */
  uint32_t blk0_mod3;    /* absolute index of the call's first block, mod 3 (noise-blanker ring slot) */
  const SdrChanCfg *cfg;
  float *state;
  unsigned long long ch_stride;
  const SdrGroup *groups;
  const float *agc_luts; /* [n_luts][132] */
  const SdrTables *tabs;
  uint32_t n_groups;
  uint32_t diag_skip;    /* diagnostics (profiling runs only): bit w set = stage w idles; results are then meaningless */
  unsigned long long *prof; /* optional [n_groups][SDR_PROF_SLOTS]: busy cycles per warp role + CTA total (diagnostics) */
  unsigned long long map_ssb, map_env; /* physical warp -> stage, 4 bits per warp (see sdr_kernel.cu) */
} SdrLaunch;

/* default placement of the 14 stages on the 14 warps of a CTA (warp id % 4 = SM sub-partition, higher id = preferred by the
 * scheduler); SDR_MAP_SSB / SDR_MAP_ENV (hex) override it for experiments */
#define SDR_MAP_SSB_DEFAULT 0x3BADC548961720ull
#define SDR_MAP_ENV_DEFAULT 0xA0D459B1328C67ull

#define SDR_STAGES 14     /* pipeline stages = warps per CTA */
#define SDR_PROF_SLOTS 512 /* [0..95] counters, [128..] a time line of four steps of every stage (diagnostics twin only) */

#define SDR_AGC_LUT_STRIDE 132

#endif
