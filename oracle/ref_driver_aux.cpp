/* oracle/ref_driver_aux.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Host harness around the UNMODIFIED neighbour blocks of the receiver chain
 *   /root/reference/SRC/AudioSDRlib/AudioSDRpreProcessor.{h,cpp}   kind 1
 *   /root/reference/SRC/AudioSDRlib/AudioIQgenerator.{h,cpp}       kind 2
 *   /root/reference/SRC/AudioSDRlib/AudioGrabberComplex256.{h,cpp} kind 3 (row 3: no arithmetic, a 256-sample snapshot)
 * #included by path (never copied) through oracle/shim/.  Built by oracle/Makefile into oracle/_ref/refaux.
 * The pre-processor's FFT comes from oracle/aux_fft128.h (CMSIS-DSP is not in the reference tree).
 * One fork()ed child per channel: AudioIQgenerator keeps its history in function statics (AudioIQgenerator.cpp:37-40).
 *
 *   refaux run   <request> <response> [jobs]
 *   refaux bench <request> <seconds>  [jobs]      time inside update() only
 *
 * Request : char magic[8]="REFAUX01"; u32 kind, n_channels, n_blocks, n_events;
 *           n_events x { u32 channel (0xFFFFFFFF = all); u32 block; u32 opcode; f32 a0 }   applied before update() of `block`
 *           kind 1: i16 I[C][S], Q[C][S]      kind 2: i16 X[C][S]                           (S = 128*n_blocks)
 * Response: char magic[8]="REFAUO01"; u32 kind, n_channels, n_blocks, n_status(=8);
 *           i16 out0[C][S], out1[C][S]  (the blocks transmitted on outputs 0 and 1);  i32 status[C][8]
 * Opcodes : kind 1: 1 startAutoI2SerrorDetection, 2 stopAutoI2SerrorDetection, 3 setI2SerrorCompensation(a0), 4 swapIQ(a0)
 *           kind 2: 1 setGainBalance(a0)
 * Status  : kind 1: {autoDetect, I2Scorrection, failureCount, successCount, savedSample, IQswap}
 *           kind 3: n_status = 520: {newDataAvailable() before the grab, _dataBufferValid, 0.., then at [8..519] the 512 int16
 *                   that grab() delivered after the last block (-1 where grab() left the destination untouched)}; inputs as
 *                   kind 1, no output planes (out0/out1 are written as zeros)
 */
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <time.h>

#define private public
#include "SRC/AudioSDRlib/AudioSDRpreProcessor.h"
#include "SRC/AudioSDRlib/AudioSDRpreProcessor.cpp"
#include "SRC/AudioSDRlib/AudioIQgenerator.h"
#include "SRC/AudioSDRlib/AudioIQgenerator.cpp"
#include "SRC/AudioSDRlib/AudioGrabberComplex256.h"
#include "SRC/AudioSDRlib/AudioGrabberComplex256.cpp"
#undef private

struct Event { uint32_t channel, block, opcode; float a0; };
struct Header { char magic[8]; uint32_t kind, n_channels, n_blocks, n_extra; };
static const int N_STATUS = 8, N_STATUS_GRAB = 520;

struct Request {
  Header h;
  std::vector<Event> events;
  const int16_t *p0, *p1;
};

static bool load(const char *path, Request &r) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); return false; }
  fseek(f, 0, SEEK_END); long len = ftell(f); fseek(f, 0, SEEK_SET);
  char *map = (char *)malloc(len);
  if (fread(map, 1, len, f) != (size_t)len) { fclose(f); return false; }
  fclose(f);
  memcpy(&r.h, map, sizeof(Header));
  if (memcmp(r.h.magic, "REFAUX01", 8) || r.h.kind < 1 || r.h.kind > 3) { fprintf(stderr, "bad request\n"); return false; }
  const char *p = map + sizeof(Header);
  r.events.resize(r.h.n_extra);
  memcpy(r.events.data(), p, sizeof(Event) * r.h.n_extra);
  p += sizeof(Event) * r.h.n_extra;
  size_t ns = (size_t)r.h.n_blocks * 128;
  r.p0 = (const int16_t *)p;
  r.p1 = r.p0 + (size_t)r.h.n_channels * ns;
  size_t need = sizeof(Header) + sizeof(Event) * r.h.n_extra + 2 * (r.h.kind != 2 ? 2 : 1) * (size_t)r.h.n_channels * ns;
  if ((size_t)len < need) { fprintf(stderr, "request truncated\n"); return false; }
  return true;
}

static std::vector<Event> events_for(const Request &r, uint32_t ch) {
  std::vector<Event> v;
  for (const Event &e : r.events) if (e.channel == ch || e.channel == 0xFFFFFFFFu) v.push_back(e);
  std::stable_sort(v.begin(), v.end(), [](const Event &a, const Event &b) { return a.block < b.block; });
  return v;
}

static void apply_pp(AudioSDRpreProcessor &p, const Event &e) {
  switch (e.opcode) {
    case 1: p.startAutoI2SerrorDetection(); break;
    case 2: p.stopAutoI2SerrorDetection(); break;
    case 3: p.setI2SerrorCompensation((int)e.a0); break;
    case 4: p.swapIQ(e.a0 != 0.0f); break;
    default: fprintf(stderr, "refaux: unknown pre-processor opcode %u\n", e.opcode); _exit(3);
  }
}
static void apply_iq(AudioIQgenerator &g, const Event &e) {
  if (e.opcode == 1) g.setGainBalance(e.a0);
  else { fprintf(stderr, "refaux: unknown generator opcode %u\n", e.opcode); _exit(3); }
}

template <class T> static T *make_zeroed() {
  void *mem = calloc(1, sizeof(T) + 64);
  return new (mem) T();
}

/* one update(); returns seconds spent inside it */
static inline double timed_update(AudioStream *s) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  s->update();
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

static void run_channel(const Request &r, uint32_t ch, int16_t *o0, int16_t *o1, int32_t *status) {
  std::vector<Event> ev = events_for(r, ch);
  size_t ei = 0, ns = (size_t)r.h.n_blocks * 128;
  static audio_block_t ba, bb;
  memset(status, 0, sizeof(int32_t) * (r.h.kind == 3 ? N_STATUS_GRAB : N_STATUS));
  if (r.h.kind == 3) {
    AudioGrabberComplex256 *g = make_zeroed<AudioGrabberComplex256>();
    const int16_t *I = r.p0 + ch * ns, *Q = r.p1 + ch * ns;
    for (uint32_t b = 0; b < r.h.n_blocks; b++) {
      memcpy(ba.data, I + (size_t)b * 128, 256);
      memcpy(bb.data, Q + (size_t)b * 128, 256);
      g->oracle_feed(0, &ba); g->oracle_feed(1, &bb);
      g->update();
    }
    memset(o0, 0, ns * 2); memset(o1, 0, ns * 2);
    status[0] = g->newDataAvailable() ? 1 : 0;
    status[1] = g->_dataBufferValid ? 1 : 0;
    int16_t dest[512];
    int16_t untouched[512];
    memset(dest, 0x55, sizeof dest); memset(untouched, 0x55, sizeof untouched);
    g->grab(dest);
    for (int i = 0; i < 512; i++) status[8 + i] = (status[1] || dest[i] != untouched[i]) ? dest[i] : -1;
    status[2] = g->newDataAvailable() ? 1 : 0;
    return;
  }
  if (r.h.kind == 1) {
    AudioSDRpreProcessor *pp = make_zeroed<AudioSDRpreProcessor>();
    const int16_t *I = r.p0 + ch * ns, *Q = r.p1 + ch * ns;
    for (uint32_t b = 0; b < r.h.n_blocks; b++) {
      while (ei < ev.size() && ev[ei].block <= b) apply_pp(*pp, ev[ei++]);
      memcpy(ba.data, I + (size_t)b * 128, 256);
      memcpy(bb.data, Q + (size_t)b * 128, 256);
      pp->oracle_feed(0, &ba); pp->oracle_feed(1, &bb); pp->oracle_clear_sent();
      pp->update();
      audio_block_t *s0 = pp->oracle_sent(0), *s1 = pp->oracle_sent(1);
      memcpy(o0 + (size_t)b * 128, s0->data, 256);
      memcpy(o1 + (size_t)b * 128, s1->data, 256);
    }
    while (ei < ev.size()) apply_pp(*pp, ev[ei++]);
    status[0] = pp->getAutoI2SerrorDetectionStatus() ? 1 : 0;
    status[1] = pp->getI2SerrorCompensation();
    status[2] = pp->failureCount;
    status[3] = pp->successCount;
    status[4] = pp->savedSample;
    status[5] = pp->IQswap ? 1 : 0;
  } else {
    AudioIQgenerator *g = make_zeroed<AudioIQgenerator>();
    const int16_t *X = r.p0 + ch * ns;
    for (uint32_t b = 0; b < r.h.n_blocks; b++) {
      while (ei < ev.size() && ev[ei].block <= b) apply_iq(*g, ev[ei++]);
      memcpy(ba.data, X + (size_t)b * 128, 256);
      g->oracle_feed(0, &ba); g->oracle_clear_sent();
      g->update();
      audio_block_t *s0 = g->oracle_sent(0), *s1 = g->oracle_sent(1);
      memcpy(o0 + (size_t)b * 128, s0->data, 256);
      memcpy(o1 + (size_t)b * 128, s1->data, 256);
    }
  }
}

static int cmd_run(const char *req, const char *resp, int jobs) {
  Request r;
  if (!load(req, r)) return 2;
  size_t ns = (size_t)r.h.n_blocks * 128, nch = r.h.n_channels;
  const int nst = r.h.kind == 3 ? N_STATUS_GRAB : N_STATUS;
  size_t out_len = sizeof(Header) + nch * ns * 4 + nch * nst * 4;
  char *out = (char *)mmap(0, out_len, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (out == MAP_FAILED) { perror("mmap"); return 2; }
  Header oh; memcpy(oh.magic, "REFAUO01", 8);
  oh.kind = r.h.kind; oh.n_channels = r.h.n_channels; oh.n_blocks = r.h.n_blocks; oh.n_extra = nst;
  memcpy(out, &oh, sizeof(oh));
  int16_t *o0 = (int16_t *)(out + sizeof(Header)), *o1 = o0 + nch * ns;
  int32_t *status = (int32_t *)(out + sizeof(Header) + nch * ns * 4);
  int running = 0, failed = 0;
  for (uint32_t ch = 0; ch < nch; ch++) {
    while (running >= jobs) { int st; wait(&st); running--; if (!WIFEXITED(st) || WEXITSTATUS(st)) failed++; }
    pid_t pid = fork();
    if (pid < 0) { perror("fork"); return 2; }
    if (pid == 0) { run_channel(r, ch, o0 + ch * ns, o1 + ch * ns, status + ch * nst); _exit(0); }
    running++;
  }
  while (running > 0) { int st; wait(&st); running--; if (!WIFEXITED(st) || WEXITSTATUS(st)) failed++; }
  if (failed) { fprintf(stderr, "refaux: %d channel workers failed\n", failed); return 2; }
  FILE *f = fopen(resp, "wb");
  if (!f) { perror(resp); return 2; }
  fwrite(out, 1, out_len, f);
  fclose(f);
  return 0;
}

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
      std::vector<Event> ev = events_for(r, ch);
      AudioSDRpreProcessor *pp = make_zeroed<AudioSDRpreProcessor>();
      AudioIQgenerator *g = make_zeroed<AudioIQgenerator>();
      for (const Event &e : ev) if (e.block == 0) { if (r.h.kind == 1) apply_pp(*pp, e); else apply_iq(*g, e); }
      static audio_block_t ba, bb;
      double acc = 0, samples = 0;
      auto t_end = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
      uint32_t b = 0;
      while (std::chrono::steady_clock::now() < t_end) {
        for (int rep = 0; rep < 64; rep++) {
          memcpy(ba.data, r.p0 + ch * ns + (size_t)b * 128, 256);
          if (r.h.kind == 1) {
            memcpy(bb.data, r.p1 + ch * ns + (size_t)b * 128, 256);
            pp->oracle_feed(0, &ba); pp->oracle_feed(1, &bb);
            acc += timed_update(pp);
          } else {
            g->oracle_feed(0, &ba);
            acc += timed_update(g);
          }
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
  printf("{\"workers\": %d, \"samples\": %.0f, \"wall_s\": %.4f, \"sps_update_only\": %.6e, \"sps_wall\": %.6e}\n",
         jobs, tot, wall, rate, tot / wall);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 4 && !strcmp(argv[1], "run")) return cmd_run(argv[2], argv[3], argc > 4 ? atoi(argv[4]) : 1);
  if (argc >= 4 && !strcmp(argv[1], "bench")) return cmd_bench(argv[2], atof(argv[3]), argc > 4 ? atoi(argv[4]) : 1);
  fprintf(stderr, "usage: refaux run <request> <response> [jobs] | refaux bench <request> <seconds> [jobs]\n");
  return 1;
}
