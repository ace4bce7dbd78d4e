/* tests/emu/aux_emu.cpp -- TEST SCAFFOLDING.  Runs the per-lane arithmetic of the pre-processor and I/Q generator kernels
 * (audiosdr_b200/csrc/sdr_aux_core.cuh, the very source nvcc compiles) on the host and compares it bit for bit with the
 * oracle (oracle/sdr_aux_oracle.c, linked in).  The kernels add only data movement around these functions; that part
 * is covered on the GPU (tests/test_gpu_aux.py).
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../audiosdr_b200/csrc/sdr_aux_core.cuh"
#include "../../audiosdr_b200/csrc/aux_tables.inc"

extern "C" int ora_aux_run(int kind, uint32_t n_channels, uint32_t n_blocks, const void *events, uint32_t n_events,
                           const int16_t *in0, const int16_t *in1, int16_t *out0, int16_t *out1, int32_t *status, int threads);

struct Ev { uint32_t channel, block, opcode; float a0; };
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }
static float tabf(const uint32_t *t, int i) { float f; memcpy(&f, &t[i], 4); return f; }

static long check_q15() {
  long bad = 0;
  for (int q = -32768; q <= 32767; q++) {
    const float want = (float)((double)(float)q / 32767.0), got = aux_q15_to_float(q);
    if (memcmp(&want, &got, 4)) bad++;
  }
  printf("q15_to_float: 65536 values, mismatches %ld\n", bad);
  return bad;
}

static int16_t ref_pcm(float v, float g) {
  const double d = (double)v * 32767.0 * (double)g;
  int32_t t = (d >= 2147483648.0 || d <= -2147483649.0 || d != d) ? (int32_t)0x80000000 : (int32_t)d;
  return (int16_t)t;
}
static long check_pcm() {
  long bad = 0, n = 0;
  const float gains[] = {1.0f, 1.1f, 1.0f / 1.1f, 0.5f, 3.0f, 70000.0f, -1.0f};
  for (int q = -32768; q <= 32767; q++)
    for (float g : gains) { const float v = aux_q15_to_float(q); n++; if (aux_to_pcm(v, g) != ref_pcm(v, g)) bad++; if (g == 1.0f && (int16_t)aux_to_pcm_unit(v) != ref_pcm(v, g)) bad++; }
  for (long i = 0; i < 20000000; i++) { /* random floats over the Hilbert output range and beyond */
    uint32_t bits = rnd() ^ (rnd() << 16);
    float v; memcpy(&v, &bits, 4);
    if (!(v == v) || fabsf(v) > 1e6f) continue;
    const float g = (i & 3) ? 1.0f : gains[(i >> 2) % 7];
    n++;
    if (aux_to_pcm(v, g) != ref_pcm(v, g)) bad++;
    if (fabsf(v) < 256.0f && (int16_t)aux_to_pcm_unit(v) != ref_pcm(v, 1.0f)) bad++;
  }
  { /* the unit-gain path skips the range test: the generator's outputs are bounded by 2 * sum|h| * (32768/32767) */
    double sum = 0;
    for (int k = 0; k < 64; k++) sum += fabs((double)tabf(AUX_IQ_HILBERT, k));
    if (2.0 * sum * 1.0001 >= 256.0) bad++;
    printf("sum|h| = %.4f (bound on |Q| = %.4f)\n", sum, 2.0 * sum * 1.0001);
  }
  printf("to_pcm: %ld values, mismatches %ld\n", n, bad);
  return bad;
}

/* I/Q generator: whole stream through the lane function with the kernel's segmenting, vs the oracle */
static long check_iq() {
  const int SEG = 4096, nb = 70, ns = nb * 128; /* 8960 samples: two full segments + a ragged third */
  std::vector<int16_t> x(ns), wi(ns), wq(ns), gi(ns), gq(ns);
  for (int i = 0; i < ns; i++) x[i] = (int16_t)((i % 977 == 0) ? (rnd() & 1 ? 32767 : -32768) : (int)(rnd() % 40001) - 20000);
  ora_aux_run(2, 1, nb, nullptr, 0, x.data(), nullptr, wi.data(), wq.data(), nullptr, 1);
  float h[64];
  for (int k = 0; k < 64; k++) h[k] = tabf(AUX_IQ_HILBERT, k);
  std::vector<float> xs(iq_pad(SEG + 256) + 32);
  for (int base = 0; base < ns; base += SEG) {
    for (size_t i = 0; i < xs.size(); i++) { uint32_t poison = 0xFFFFFFFFu; memcpy(&xs[i], &poison, 4); }
    for (int q = 0; q < SEG + 256; q++) {
      const long n = (long)base - 256 + q;
      xs[iq_pad(q)] = (n < 0 || n >= ns) ? 0.0f : aux_q15_to_float(x[n]);
    }
    for (int w = 0; w < 8; w++)
      for (int lane = 0; lane < 32; lane++) {
        const int q0 = 256 + 512 * w + 16 * lane, n0 = base + 512 * w + 16 * lane;
        if (n0 >= ns) continue;
        float acc[16];
        iq_lane_fir(xs.data(), q0, h, acc);
        for (int c = 0; c < 16; c++) {
          gq[n0 + c] = (int16_t)aux_to_pcm(acc[c], 1.0f);
          gi[n0 + c] = (int16_t)aux_to_pcm(iq_lane_delayed(xs.data(), q0, c), 1.0f);
        }
      }
  }
  long bad = 0;
  for (int i = 0; i < ns; i++) bad += (gi[i] != wi[i]) + (gq[i] != wq[i]);
  printf("iq generator: %d samples, mismatches %ld\n", ns, bad);
  return bad;
}

/* pre-processor: detector path block by block and feed-forward path chunk by chunk, vs the oracle */
static long check_pp() {
  const int nch = 6, nb = 90, ns = nb * 128;
  std::vector<int16_t> I((size_t)nch * ns), Q((size_t)nch * ns), wi(I.size()), wq(I.size()), gi(I.size()), gq(I.size());
  for (int c = 0; c < nch; c++)
    for (int i = 0; i < ns; i++) {
      const double ph = 2.0 * 3.14159265358979323846 * (3000.0 + 700.0 * c) / 44100.0;
      const int lag = (c % 3 == 1) ? 1 : 0, lagi = (c % 3 == 2) ? 1 : 0; /* Q or I one sample late */
      I[(size_t)c * ns + i] = (int16_t)lrint(9000.0 * cos(ph * (i - lagi)) + (int)(rnd() % 41) - 20);
      Q[(size_t)c * ns + i] = (int16_t)lrint(9000.0 * sin(ph * (i - lag)) + (int)(rnd() % 41) - 20);
    }
  std::vector<Ev> ev = {{0xFFFFFFFFu, 0, 1, 0.f}, {3, 40, 4, 1.f}, {4, 50, 3, -1.f}, {5, 60, 2, 0.f}, {0, 70, 3, 1.f}};
  std::vector<int32_t> wst((size_t)nch * 8);
  ora_aux_run(1, nch, nb, ev.data(), (uint32_t)ev.size(), I.data(), Q.data(), wi.data(), wq.data(), wst.data(), 1);
  float tw[128];
  for (int k = 0; k < 128; k++) tw[k] = tabf(AUX_FFT_TW, k);
  long bad = 0;
  for (int c = 0; c < nch; c++) {
    PpState s; memset(&s, 0, sizeof s);
    for (int b = 0; b < nb; b++) {
      for (const Ev &e : ev) if ((e.channel == (uint32_t)c || e.channel == 0xFFFFFFFFu) && e.block == (uint32_t)b) pp_apply(s, e.opcode, (int32_t)e.a0);
      const int16_t *bi = &I[(size_t)c * ns + b * 128], *bq = &Q[(size_t)c * ns + b * 128];
      int16_t *oi = &gi[(size_t)c * ns + b * 128], *oq = &gq[(size_t)c * ns + b * 128];
      if (s.autod) { /* detector path: the sequential form */
        int16_t ri[128], rq[128];
        memcpy(ri, bi, 256); memcpy(rq, bq, 256);
        pp_correct_block(ri, rq, s);
        float buf[256 * 3];
        for (int i = 0; i < 128; i++) { buf[(2 * i) * 3] = aux_q15_to_float(ri[i]); buf[(2 * i + 1) * 3] = aux_q15_to_float(rq[i]); }
        pp_fft128(buf, 3, tw);
        pp_detect(buf, 3, s);
        for (int i = 0; i < 128; i++) { oi[i] = s.swap ? rq[i] : ri[i]; oq[i] = s.swap ? ri[i] : rq[i]; }
      } else { /* feed-forward form, one call = one block here, reading the ORIGINAL planes around it */
        const int16_t *ci = &I[(size_t)c * ns], *cq = &Q[(size_t)c * ns];
        for (int n = 0; n < 128; n += 8) {
          /* emulate "call starts at this block": position relative to the call start */
          const int16_t *pi = ci + b * 128, *pq = cq + b * 128;
          const int16_t prev = (int16_t)s.saved; /* savedSample holds I's or Q's last sample, whichever rail is being delayed */
          pp_static_chunk(pi + n, pq + n, n ? pi[n - 1] : prev, n ? pq[n - 1] : prev, (n & 127) == 0, s.corr, s.swap, oi + n, oq + n);
        }
        if (s.corr == 1) s.saved = bi[127];
        else if (s.corr == -1) s.saved = bq[127];
      }
    }
    for (const Ev &e : ev) if ((e.channel == (uint32_t)c || e.channel == 0xFFFFFFFFu) && e.block >= (uint32_t)nb) pp_apply(s, e.opcode, (int32_t)e.a0);
    const int32_t got[6] = {s.autod, s.corr, s.fail, s.succ, s.saved, s.swap};
    for (int k = 0; k < 6; k++) if (got[k] != wst[(size_t)c * 8 + k]) { bad++; printf("  channel %d status %d: %d vs %d\n", c, k, got[k], wst[(size_t)c * 8 + k]); }
  }
  for (size_t i = 0; i < I.size(); i++) bad += (gi[i] != wi[i]) + (gq[i] != wq[i]);
  /* feed-forward form over a multi-block call (chunks that cross block boundaries), all three corrections, with swap */
  for (int corr = -1; corr <= 1; corr++)
    for (int swap = 0; swap <= 1; swap++) {
      std::vector<Ev> e2 = {{0xFFFFFFFFu, 0, 3, (float)corr}, {0xFFFFFFFFu, 0, 4, (float)swap}};
      ora_aux_run(1, 1, nb, e2.data(), 2, I.data(), Q.data(), wi.data(), wq.data(), nullptr, 1);
      for (int n = 0; n < ns; n += 8)
        pp_static_chunk(&I[n], &Q[n], n ? I[n - 1] : (int16_t)0, n ? Q[n - 1] : (int16_t)0, (n & 127) == 0, corr, swap, &gi[n], &gq[n]);
      for (int i = 0; i < ns; i++) bad += (gi[i] != wi[i]) + (gq[i] != wq[i]);
    }
  printf("pre-processor: %d channels x %d blocks (+6 feed-forward runs), mismatches %ld\n", nch, nb, bad);
  return bad;
}

int main() {
  long bad = check_q15() + check_pcm() + check_iq() + check_pp();
  printf("total mismatches %ld\n", bad);
  return bad ? 1 : 0;
}
