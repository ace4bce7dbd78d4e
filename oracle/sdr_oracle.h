/* oracle/sdr_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference receiver chain AudioSDR::update()
 * (/root/reference/SRC/AudioSDRlib/AudioSDR.cpp:39-168 and the helpers it calls),
 * one explicit state object per channel (the reference's function-static state,
 * AudioSDR.cpp:41-44 and 690-694, becomes per-channel members: SURVEY.md Q1/Q2).
 *
 * PARITY PINNING: the reference ships no tests or golden vectors.  This
 * restatement is pinned bit-for-bit against the unmodified reference compiled on
 * the host (oracle/_ref/refsdr, built by oracle/Makefile) by tests/test_oracle_vs_ref.py
 * and against the fixtures that binary generated (tests/golden/, tools/gen_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use
 * this library; the product (audiosdr_b200/) never links or loads it.
 *
 * Citation tags:  C: = SRC/AudioSDRlib/AudioSDR.cpp   H: = SRC/AudioSDRlib/AudioSDR.h
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORA_NBLOCK 128
#define ORA_NSTATUS 16

/* per-update stage taps (each 128 floats), in chain order */
enum {
  ORA_TAP_IN_I = 0, ORA_TAP_IN_Q, /* after input scaling          C:67-70   */
  ORA_TAP_NB_I, ORA_TAP_NB_Q,     /* after noise blanker          C:73      */
  ORA_TAP_IF_I, ORA_TAP_IF_Q,     /* after IF band-pass           C:77-78   */
  ORA_TAP_DM_I, ORA_TAP_DM_Q,     /* _Idata/_Qdata when the demodulator is done */
  ORA_TAP_DEMOD,                  /* _audioOut after demodulation C:84-144  */
  ORA_TAP_AUDF,                   /* after audio band-pass        C:149     */
  ORA_TAP_AGC,                    /* after AGC                    C:152     */
  ORA_TAP_ALS,                    /* after ALS                    C:155     */
  ORA_NTAPS
};

typedef struct { uint32_t channel, block, opcode; float a0, a1, a2; } ora_event;

typedef struct ora_channel ora_channel;

ora_channel *ora_new(void);                       /* zeroed storage + constructor defaults + init()  (H:77-79, C:174-185) */
void ora_free(ora_channel *c);
int ora_apply(ora_channel *c, uint32_t opcode, float a0, float a1, float a2); /* setter surface, opcodes = oracle/ref_client.py OPS */
void ora_set_taps(ora_channel *c, float *taps);   /* NULL or ORA_NTAPS*128 floats, rewritten by every update */
/* One 128-sample block, int16 wire format (the reference's audio_block_t boundary). */
void ora_update_i16(ora_channel *c, const int16_t *I, const int16_t *Q, float *audio, int16_t *pcm);
/* Same, but the block arrives as float32 x (e.g. q/32767): scaled sample = (float)((double)x * gain). */
void ora_update_f32(ora_channel *c, const float *I, const float *Q, float *audio, int16_t *pcm);
void ora_status(const ora_channel *c, float *status /* ORA_NSTATUS */);
/* debugging access to the NCO phases (function-statics in the reference) */
void ora_get_phases(const ora_channel *c, float *phase_ssb, float *phase_am);

/* Batch helper: every channel is an independent object; events as in oracle/ref_driver.cpp.
 * I/Q are [n_channels][n_blocks*128] (int16 when fmt==0, float32 when fmt==1); audio/pcm/status may be NULL.
 * n_threads>1 splits channels over pthreads.  Returns seconds spent inside the update loop (max over threads). */
double ora_run(uint32_t n_channels, uint32_t n_blocks, const ora_event *events, uint32_t n_events, int fmt,
               const void *I, const void *Q, float *audio, int16_t *pcm, float *status, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
