/* Test-infrastructure shim (NOT product code): stand-in for the Teensy Audio
 * Library runtime that the reference sits behind (reference H:75-79,163;
 * C:46-56,164-167).  Sample rate is the Teensy-4 float literal 44100.0f
 * (SURVEY.md 8c); block size 128. */
#ifndef ORACLE_SHIM_AUDIOSTREAM_H
#define ORACLE_SHIM_AUDIOSTREAM_H
#include <stdint.h>
#define AUDIO_BLOCK_SAMPLES 128
#define AUDIO_SAMPLE_RATE_EXACT 44100.0f
#define AUDIO_SAMPLE_RATE AUDIO_SAMPLE_RATE_EXACT
typedef struct audio_block_struct {
  uint8_t ref_count;
  uint8_t reserved1;
  uint16_t memory_pool_index;
  int16_t data[AUDIO_BLOCK_SAMPLES];
} audio_block_t;

class AudioStream {
 public:
  AudioStream(unsigned char ninput, audio_block_t **iqueue)
      : num_inputs(ninput), inputQueue(iqueue) {
    for (int i = 0; i < ninput; i++) inputQueue[i] = 0;
    for (int i = 0; i < 4; i++) sent[i] = 0;
  }
  virtual ~AudioStream() {}
  virtual void update(void) = 0;
  /* harness side */
  void oracle_feed(unsigned idx, audio_block_t *b) { inputQueue[idx] = b; }
  audio_block_t *oracle_sent(unsigned idx) { return sent[idx]; }
  void oracle_clear_sent() { for (int i = 0; i < 4; i++) sent[i] = 0; }
 protected:
  audio_block_t *receiveWritable(unsigned int index = 0) {
    if (index >= num_inputs) return 0;
    audio_block_t *b = inputQueue[index];
    inputQueue[index] = 0;
    return b;
  }
  audio_block_t *receiveReadOnly(unsigned int index = 0) { return receiveWritable(index); }
  /* AudioSDR never allocates; AudioIQgenerator allocates its Q output block every update (AudioIQgenerator.cpp:47) */
  static audio_block_t *allocate(void) { static audio_block_t pool[4]; static unsigned k = 0; return &pool[k++ & 3]; }
  void transmit(audio_block_t *block, unsigned char index = 0) { if (index < 4) sent[index] = block; }
  static void release(audio_block_t *) {}
  unsigned char num_inputs;
 private:
  audio_block_t **inputQueue;
  audio_block_t *sent[4];
};
#endif
