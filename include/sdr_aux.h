/* include/sdr_aux.h -- C ABI of the two blocks either side of the receiver chain, batched over
 * independent channels on one GPU.  Library: audiosdr_b200/libsdr_aux.so (CUDA, sm_100a; no CPU fallback).
 *
 *   sdr_preproc_*   replaces  class AudioSDRpreProcessor   AudioSDRpreProcessor.h:49-73, update() AudioSDRpreProcessor.cpp:46-138
 *   sdr_iqgen_*     replaces  class AudioIQgenerator       AudioIQgenerator.h:49-107,   update() AudioIQgenerator.cpp:33-87
 *   sdr_grabber_*   replaces  class AudioGrabberComplex256 AudioGrabberComplex256.h:46-64, update()/grab() AudioGrabberComplex256.cpp:50-91
 *
 * Planes are int16 (the reference's audio_block_t wire format), channel-major: sample s of channel c is plane[c*pitch + s],
 * pitch in elements.  Base pointers must be 16-byte aligned and pitches multiples of 8.  Output planes must not overlap the
 * input planes (the reference works in place on one block; a batched one-sample shift cannot).  Each process call advances
 * every channel by n_blocks blocks of 128 samples; all per-channel state lives in device memory between calls.
 * All functions return 0 (SDR_AUX_OK) or a negative error; sdr_aux_last_error() gives the text (thread local).
 */
#ifndef SDR_AUX_H
#define SDR_AUX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { SDR_AUX_OK = 0, SDR_AUX_EINVAL = -1, SDR_AUX_ECUDA = -2, SDR_AUX_ENOMEM = -3 };

/* ---------------------------------------------------------------- pre-processor ---- */
typedef struct sdr_preproc sdr_preproc_t;

/* setter ids = the reference's public functions (AudioSDRpreProcessor.h:54-59) */
typedef enum {
  SDR_PP_startAutoI2SerrorDetection = 1, /* AudioSDRpreProcessor.cpp:142-148 */
  SDR_PP_stopAutoI2SerrorDetection = 2,  /* AudioSDRpreProcessor.cpp:151-154 */
  SDR_PP_setI2SerrorCompensation = 3,    /* arg = -1, 0, +1;  AudioSDRpreProcessor.cpp:160-163 */
  SDR_PP_swapIQ = 4                      /* arg = 0 / 1;      AudioSDRpreProcessor.cpp:169 */
} sdr_preproc_setter;

typedef struct {
  int32_t auto_detect;   /* getAutoI2SerrorDetectionStatus(), AudioSDRpreProcessor.cpp:157 */
  int32_t correction;    /* getI2SerrorCompensation(),        AudioSDRpreProcessor.cpp:166 */
  int32_t failure_count; /* private members, for tests and checkpoints */
  int32_t success_count;
  int32_t saved_sample;
  int32_t swap;
} sdr_preproc_status;

int sdr_preproc_create(sdr_preproc_t **out, uint32_t n_channels, int device);
void sdr_preproc_destroy(sdr_preproc_t *h);
/* channels == NULL: every channel.  Setters take effect before the next processed block, in call order. */
int sdr_preproc_set(sdr_preproc_t *h, const uint32_t *channels, uint32_t n, uint32_t setter, int32_t arg);
int sdr_preproc_get_status(sdr_preproc_t *h, const uint32_t *channels, uint32_t n, sdr_preproc_status *out);
/* AudioSDRpreProcessor::update() for every channel, n_blocks times.  Device planes; asynchronous on `cuda_stream`. */
int sdr_preproc_process_device(sdr_preproc_t *h, const int16_t *I, const int16_t *Q, size_t in_pitch, int16_t *I_out,
                               int16_t *Q_out, size_t out_pitch, uint32_t n_blocks, void *cuda_stream);
/* ... host planes: copies in, runs, copies out, returns when the outputs are in host memory */
int sdr_preproc_process_host(sdr_preproc_t *h, const int16_t *I, const int16_t *Q, size_t in_pitch, int16_t *I_out,
                             int16_t *Q_out, size_t out_pitch, uint32_t n_blocks);
uint64_t sdr_preproc_launch_count(const sdr_preproc_t *h);

/* ---------------------------------------------------------------- I/Q generator ---- */
typedef struct sdr_iqgen sdr_iqgen_t;

int sdr_iqgen_create(sdr_iqgen_t **out, uint32_t n_channels, int device);
void sdr_iqgen_destroy(sdr_iqgen_t *h);
/* AudioIQgenerator::setGainBalance (AudioIQgenerator.h:56-60): gainI = balance, gainQ = 1.0 / balance */
int sdr_iqgen_set_gain_balance(sdr_iqgen_t *h, const uint32_t *channels, uint32_t n, float balance);
/* AudioIQgenerator::update() for every channel, n_blocks times: X real input -> I (delayed 128 samples) and Q (Hilbert) */
int sdr_iqgen_process_device(sdr_iqgen_t *h, const int16_t *X, size_t in_pitch, int16_t *I_out, int16_t *Q_out,
                             size_t out_pitch, uint32_t n_blocks, void *cuda_stream);
int sdr_iqgen_process_host(sdr_iqgen_t *h, const int16_t *X, size_t in_pitch, int16_t *I_out, int16_t *Q_out,
                           size_t out_pitch, uint32_t n_blocks);
uint64_t sdr_iqgen_launch_count(const sdr_iqgen_t *h);

/* ---------------------------------------------------------------- complex snapshot grabber ---- */
/* replaces class AudioGrabberComplex256 (AudioGrabberComplex256.h:46-64): every update() appends one block of interleaved
 * (re, im) samples to a two-block buffer; each time the buffer fills, it becomes the snapshot grab() hands out
 * (AudioGrabberComplex256.cpp:50-91).  No arithmetic: the batched form is a gather of the last complete pair of blocks. */
typedef struct sdr_grabber sdr_grabber_t;

int sdr_grabber_create(sdr_grabber_t **out, uint32_t n_channels, int device);
void sdr_grabber_destroy(sdr_grabber_t *h);
/* AudioGrabberComplex256::update() for every channel, n_blocks times; device planes as for the pre-processor */
int sdr_grabber_process_device(sdr_grabber_t *h, const int16_t *I, const int16_t *Q, size_t pitch, uint32_t n_blocks, void *cuda_stream);
/* newDataAvailable() of one channel: 1 / 0, negative on error */
int sdr_grabber_new_data_available(sdr_grabber_t *h, uint32_t channel);
/* grab() for the listed channels (NULL: all): 512 int16 per channel = 256 complex samples, to HOST memory
 * dest[i*512 .. i*512+511].  Before the first complete pair nothing is written (the reference's _dataBufferValid), and the
 * call still clears the channels' new-data flags.  Returns the number of channels written, negative on error. */
int sdr_grabber_grab(sdr_grabber_t *h, const uint32_t *channels, uint32_t n, int16_t *dest);
/* all channels, device to device: dest[n_channels][512]; returns 0 when nothing was valid yet, 1 when written */
int sdr_grabber_grab_device(sdr_grabber_t *h, int16_t *dest, void *cuda_stream);

/* Spectrum tap on the snapshots -- what the consumer of grab() does with them (the sketch's panadapter / S-meter; not part
 * of the library classes): the 256-point forward complex FFT of a channel's snapshot, samples taken as
 * floats (re, im) = (float)int16, and the power re^2 + im^2 per bin, natural bin order (bin k = k * 44100/256 Hz, bins 128..255
 * = negative frequencies).  power[i*256 + k] for the i-th listed channel (NULL: all channels).  Returns the number of
 * channels written (host form) / 1 (device form), 0 while no snapshot is valid yet, negative on error; the new-data flags
 * are left alone (a spectrum is a view of the snapshot, not a grab()). */
int sdr_grabber_spectrum(sdr_grabber_t *h, const uint32_t *channels, uint32_t n, float *power);
int sdr_grabber_spectrum_device(sdr_grabber_t *h, const uint32_t *d_channels, uint32_t n, float *d_power, void *cuda_stream);

const char *sdr_aux_last_error(void);
const char *sdr_aux_version(void);

#ifdef __cplusplus
}
#endif
#endif
