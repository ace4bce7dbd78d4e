/* oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Host harness around the UNMODIFIED reference receiver
 *   /root/reference/SRC/AudioSDRlib/AudioSDR.{h,cpp}
 * which is #included by path (never copied) through the header shim in
 * oracle/shim/.  Built by oracle/Makefile into oracle/_ref/refsdr.
 *
 * Why a separate process per channel: the reference keeps the Hilbert history
 * rings, both NCO phases (AudioSDR.cpp:41-44) and the SAM oscillator feedback
 * state (AudioSDR.cpp:690-694) in function-static variables, so two instances
 * in one address space would share them.  Every channel therefore runs in a
 * fork()ed child whose statics are pristine.
 * Why zeroed storage: several members read by update() are never initialised
 * (AudioSDR.h:210-229,270,274,325-326); the shipped sketch uses a global
 * (BareBonesWSPR.ino:52), i.e. static-storage zero semantics.
 *
 *   refsdr run   <request> <response> [jobs]   golden output for every channel
 *   refsdr bench <request> <seconds>  [jobs]   time update() only, one worker per job
 *
 * Request file  (little endian):
 *   char magic[8] = "REFSDR01"; u32 n_channels, n_blocks, n_events, reserved;
 *   n_events x { u32 channel (0xFFFFFFFF = all); u32 block; u32 opcode; f32 a0,a1,a2 }
 *       -- applied in file order, immediately before update() of `block`
 *   i16 I[n_channels][n_blocks*128];  i16 Q[n_channels][n_blocks*128];
 * Response file:
 *   char magic[8] = "REFOUT01"; u32 n_channels, n_blocks, n_status, reserved;
 *   f32 audio[n_channels][n_blocks*128]   = muted ? 0 : _output_gain*_audioOut   (float product, C:160)
 *   i16 pcm  [n_channels][n_blocks*128]   = the int16 block the reference transmits (C:158-165)
 *   f32 status[n_channels][n_status]      = getters after the last block
 */
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <time.h>

#define private public /* expose _audioOut/_output_gain for the f32 tap; after the STL includes */
#include "SRC/AudioSDRlib/AudioSDR.h"
#include "SRC/AudioSDRlib/AudioSDR.cpp"
#undef private

enum {
  OP_setMute = 1, OP_setInputGain, OP_setIQgainBalance, OP_setDemodMode, OP_enableAudioFilter,
  OP_disableAudioFilter, OP_setOutputGain, OP_setAudioFilter, OP_enableALSfilter, OP_disableALSfilter,
  OP_setALSfilterNotch, OP_setALSfilterPeak, OP_setALSfilterAdaptive, OP_setALSfilterStatic,
  OP_setALSfilterParams, OP_enableAGC, OP_disableAGC, OP_setAGCthreshold, OP_setAGCslope, OP_setAGCmode,
  OP_setAGCkneeWidth, OP_setAGCattackTime, OP_setAGCreleaseTime, OP_setAGChangTime, OP_setAGCstaticGain,
  OP_enableNoiseBlanker, OP_disableNoiseBlanker, OP_setNoiseBlankerThreshold, OP_setNoiseBlankerThresholdDb,
  OP_init = 30,
  OP_oracle_identity_IF = 100 /* harness-only: point the IF cascades at {1,0,0,0,0}x4 (SURVEY 8c stage taps) */
};

struct Event { uint32_t channel, block, opcode; float a0, a1, a2; };
struct Header { char magic[8]; uint32_t n_channels, n_blocks, n_extra, reserved; };
static const int N_STATUS = 16;
static float identity_coefs[20] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0};

static void apply(AudioSDR &s, const Event &e) {
  switch (e.opcode) {
    case OP_setMute: s.setMute(e.a0 != 0.0f); break;
    case OP_setInputGain: s.setInputGain(e.a0); break;
    case OP_setIQgainBalance: s.setIQgainBalance(e.a0); break;
    case OP_setDemodMode: s.setDemodMode((int)e.a0); break;
    case OP_enableAudioFilter: s.enableAudioFilter(); break;
    case OP_disableAudioFilter: s.disableAudioFilter(); break;
    case OP_setOutputGain: s.setOutputGain(e.a0); break;
    case OP_setAudioFilter: s.setAudioFilter((int)e.a0); break;
    case OP_enableALSfilter: s.enableALSfilter(); break;
    case OP_disableALSfilter: s.disableALSfilter(); break;
    case OP_setALSfilterNotch: s.setALSfilterNotch(); break;
    case OP_setALSfilterPeak: s.setALSfilterPeak(); break;
    case OP_setALSfilterAdaptive: s.setALSfilterAdaptive(); break;
    case OP_setALSfilterStatic: s.setALSfilterStatic(); break;
    case OP_setALSfilterParams: s.setALSfilterParams((unsigned)e.a0, e.a1, e.a2); break;
    case OP_enableAGC: s.enableAGC(); break;
    case OP_disableAGC: s.disableAGC(); break;
    case OP_setAGCthreshold: s.setAGCthreshold(e.a0); break;
    case OP_setAGCslope: s.setAGCslope(e.a0); break;
    case OP_setAGCmode: s.setAGCmode((int16_t)e.a0); break;
    case OP_setAGCkneeWidth: s.setAGCkneeWidth(e.a0); break;
    case OP_setAGCattackTime: s.setAGCattackTime(e.a0); break;
    case OP_setAGCreleaseTime: s.setAGCreleaseTime(e.a0); break;
    case OP_setAGChangTime: s.setAGChangTime(e.a0); break;
    case OP_setAGCstaticGain: s.setAGCstaticGain(e.a0); break;
    case OP_enableNoiseBlanker: s.enableNoiseBlanker(); break;
    case OP_disableNoiseBlanker: s.disableNoiseBlanker(); break;
    case OP_setNoiseBlankerThreshold: s.setNoiseBlankerThreshold(e.a0); break;
    case OP_setNoiseBlankerThresholdDb: s.setNoiseBlankerThresholdDb(e.a0); break;
    case OP_init: s.init(); break;
    case OP_oracle_identity_IF:
      s._IFfilterI.pCoeffs = identity_coefs;
      s._IFfilterQ.pCoeffs = identity_coefs;
      break;
    default: fprintf(stderr, "refsdr: unknown opcode %u\n", e.opcode); _exit(3);
  }
}

static AudioSDR *make_instance() {
  void *mem = calloc(1, sizeof(AudioSDR) + 64);
  return new (mem) AudioSDR();
}

struct Request {
  Header h;
  std::vector<Event> events;
  const int16_t *I, *Q;
  void *map; size_t map_len;
};

static bool load(const char *path, Request &r) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); return false; }
  fseek(f, 0, SEEK_END); long len = ftell(f); fseek(f, 0, SEEK_SET);
  r.map = malloc(len); r.map_len = len;
  if (fread(r.map, 1, len, f) != (size_t)len) { fclose(f); return false; }
  fclose(f);
  memcpy(&r.h, r.map, sizeof(Header));
  if (memcmp(r.h.magic, "REFSDR01", 8)) { fprintf(stderr, "bad magic\n"); return false; }
  const char *p = (const char *)r.map + sizeof(Header);
  r.events.resize(r.h.n_extra);
  memcpy(r.events.data(), p, sizeof(Event) * r.h.n_extra);
  p += sizeof(Event) * r.h.n_extra;
  size_t ns = (size_t)r.h.n_blocks * 128;
  r.I = (const int16_t *)p;
  r.Q = r.I + (size_t)r.h.n_channels * ns;
  size_t need = sizeof(Header) + sizeof(Event) * r.h.n_extra + 4 * (size_t)r.h.n_channels * ns;
  if ((size_t)len < need) { fprintf(stderr, "request truncated\n"); return false; }
  return true;
}

/* Per-channel event list, ordered by block (stable in file order). */
static std::vector<Event> events_for(const Request &r, uint32_t ch) {
  std::vector<Event> v;
  for (const Event &e : r.events) if (e.channel == ch || e.channel == 0xFFFFFFFFu) v.push_back(e);
  std::stable_sort(v.begin(), v.end(), [](const Event &a, const Event &b) { return a.block < b.block; });
  return v;
}

static void run_channel(const Request &r, uint32_t ch, float *audio, int16_t *pcm, float *status) {
  AudioSDR *sdr = make_instance();
  std::vector<Event> ev = events_for(r, ch);
  size_t ei = 0;
  size_t ns = (size_t)r.h.n_blocks * 128;
  const int16_t *I = r.I + ch * ns, *Q = r.Q + ch * ns;
  static audio_block_t bi, bq;
  for (uint32_t b = 0; b < r.h.n_blocks; b++) {
    while (ei < ev.size() && ev[ei].block <= b) apply(*sdr, ev[ei++]);
    memcpy(bi.data, I + (size_t)b * 128, 256);
    memcpy(bq.data, Q + (size_t)b * 128, 256);
    sdr->oracle_feed(0, &bi);
    sdr->oracle_feed(1, &bq);
    sdr->oracle_clear_sent();
    sdr->update();
    audio_block_t *o = sdr->oracle_sent(0);
    for (int i = 0; i < 128; i++) {
      pcm[(size_t)b * 128 + i] = o ? o->data[i] : 0;
      audio[(size_t)b * 128 + i] = sdr->_isMuted ? 0.0f : sdr->_output_gain * sdr->_audioOut[i];
    }
  }
  while (ei < ev.size()) apply(*sdr, ev[ei++]);
  memset(status, 0, sizeof(float) * N_STATUS);
  status[0] = sdr->getTuningOffset();
  status[1] = (float)sdr->getDemodMode();
  status[2] = sdr->AGCisActive() ? 1.0f : 0.0f;
  status[3] = sdr->NoiseBlankerDetection() ? 1.0f : 0.0f;
  status[4] = sdr->getSAMfrequency();
  status[5] = sdr->getSAMphaseLockStatus() ? 1.0f : 0.0f;
  status[6] = sdr->getAMcarrierLevel();
  status[7] = sdr->getBPFlower();
  status[8] = sdr->getBPFupper();
  status[9] = sdr->getMute() ? 1.0f : 0.0f;
  status[10] = (float)sdr->getAudioFilter();
  status[11] = sdr->AGCisEnabled() ? 1.0f : 0.0f;
  status[12] = sdr->NoiseBlankerisEnabled() ? 1.0f : 0.0f;
  status[13] = sdr->ALSfilterIsEnabled() ? 1.0f : 0.0f;
  status[14] = sdr->_agc_gain;
  status[15] = sdr->_nb_AvgMag;
}

static int cmd_run(const char *req, const char *resp, int jobs) {
  Request r;
  if (!load(req, r)) return 2;
  size_t ns = (size_t)r.h.n_blocks * 128;
  size_t nch = r.h.n_channels;
  size_t out_len = sizeof(Header) + nch * ns * 4 + nch * ns * 2 + nch * N_STATUS * 4;
  char *out = (char *)mmap(0, out_len, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (out == MAP_FAILED) { perror("mmap"); return 2; }
  Header oh; memcpy(oh.magic, "REFOUT01", 8);
  oh.n_channels = r.h.n_channels; oh.n_blocks = r.h.n_blocks; oh.n_extra = N_STATUS; oh.reserved = 0;
  memcpy(out, &oh, sizeof(oh));
  float *audio = (float *)(out + sizeof(Header));
  int16_t *pcm = (int16_t *)(out + sizeof(Header) + nch * ns * 4);
  float *status = (float *)(out + sizeof(Header) + nch * ns * 6);
  int running = 0, failed = 0;
  for (uint32_t ch = 0; ch < nch; ch++) {
    while (running >= jobs) { int st; wait(&st); running--; if (!WIFEXITED(st) || WEXITSTATUS(st)) failed++; }
    pid_t pid = fork();
    if (pid < 0) { perror("fork"); return 2; }
    if (pid == 0) {
      run_channel(r, ch, audio + ch * ns, pcm + ch * ns, status + ch * N_STATUS);
      _exit(0);
    }
    running++;
  }
  while (running > 0) { int st; wait(&st); running--; if (!WIFEXITED(st) || WEXITSTATUS(st)) failed++; }
  if (failed) { fprintf(stderr, "refsdr: %d channel workers failed\n", failed); return 2; }
  FILE *f = fopen(resp, "wb");
  if (!f) { perror(resp); return 2; }
  fwrite(out, 1, out_len, f);
  fclose(f);
  return 0;
}

/* bench: each worker owns one channel of the request and streams its blocks
 * round-robin (state carries on, the signal simply repeats) until `seconds`
 * of wall time elapsed; only the time spent inside update() is accumulated. */
static int cmd_bench(const char *req, double seconds, int jobs) {
  Request r;
  if (!load(req, r)) return 2;
  struct Slot { double upd_s; double samples; };
  Slot *slots = (Slot *)mmap(0, sizeof(Slot) * jobs, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  size_t ns = (size_t)r.h.n_blocks * 128;
  auto wall0 = std::chrono::steady_clock::now();
  for (int w = 0; w < jobs; w++) {
    pid_t pid = fork();
    if (pid == 0) {
      uint32_t ch = (uint32_t)(w % r.h.n_channels);
      AudioSDR *sdr = make_instance();
      std::vector<Event> ev = events_for(r, ch);
      for (const Event &e : ev) if (e.block == 0) apply(*sdr, e);
      const int16_t *I = r.I + ch * ns, *Q = r.Q + ch * ns;
      static audio_block_t bi, bq;
      double acc = 0, samples = 0;
      auto t_end = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
      uint32_t b = 0;
      while (std::chrono::steady_clock::now() < t_end) {
        for (int rep = 0; rep < 64; rep++) {
          memcpy(bi.data, I + (size_t)b * 128, 256);
          memcpy(bq.data, Q + (size_t)b * 128, 256);
          sdr->oracle_feed(0, &bi);
          sdr->oracle_feed(1, &bq);
          struct timespec t0, t1;
          clock_gettime(CLOCK_MONOTONIC, &t0);
          sdr->update();
          clock_gettime(CLOCK_MONOTONIC, &t1);
          acc += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
          samples += 128;
          if (++b == r.h.n_blocks) b = 0;
        }
      }
      slots[w].upd_s = acc; slots[w].samples = samples;
      _exit(0);
    }
  }
  for (int w = 0; w < jobs; w++) { int st; wait(&st); }
  double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
  double tot = 0, rate = 0;
  for (int w = 0; w < jobs; w++) { tot += slots[w].samples; rate += slots[w].samples / slots[w].upd_s; }
  /* aggregate = sum over workers of (samples / time inside update()); wall-clock rate also given */
  printf("{\"workers\": %d, \"samples\": %.0f, \"wall_s\": %.4f, \"sps_update_only\": %.6e, \"sps_wall\": %.6e}\n",
         jobs, tot, wall, rate, tot / wall);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 4 && !strcmp(argv[1], "run")) return cmd_run(argv[2], argv[3], argc > 4 ? atoi(argv[4]) : 1);
  if (argc >= 4 && !strcmp(argv[1], "bench")) return cmd_bench(argv[2], atof(argv[3]), argc > 4 ? atoi(argv[4]) : 1);
  fprintf(stderr, "usage: refsdr run <request> <response> [jobs] | refsdr bench <request> <seconds> [jobs]\n");
  return 1;
}
