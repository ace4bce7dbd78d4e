/* tests/emu/exhaustive_lut.cpp -- TEST SCAFFOLDING.
 * Exhaustive proof that the product's divide-free sine-table index (sdrk::lut_index, sdr_pipeline.cuh) equals the
 * reference expression (long)(Phase * 65535.0 / twoPI) of AudioSDR.h:364 for EVERY float Phase in [0, 8),
 * i.e. all 2^30 + 2^23*... bit patterns 0x00000000 .. 0x40FFFFFF (the oscillator only ever sees [0, 2*pi + pi/2]). */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include "../../audiosdr_b200/csrc/sdr_pipeline.cuh"

int main() {
  const uint32_t END = 0x41000000u; /* 8.0f */
  const float two_pi = (float)(2.0 * SDR_PI_D);
  unsigned nt = std::thread::hardware_concurrency(); if (!nt) nt = 4;
  std::atomic<uint64_t> bad(0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t]() {
      uint64_t lo = (uint64_t)END * t / nt, hi = (uint64_t)END * (t + 1) / nt, b = 0;
      for (uint64_t u = lo; u < hi; u++) {
        uint32_t bits = (uint32_t)u; float ph; memcpy(&ph, &bits, 4);
        int want = (int)((uint16_t)(long)((double)ph * 65535.0 / (double)two_pi));
        int got = sdrk::lut_index(ph);
        if (ph < two_pi && want != got) { if (b < 5) fprintf(stderr, "mismatch at %a: want %d got %d\n", ph, want, got); b++; }
        if (sdrk::lut_index_lt8(ph) != got) { if (b < 5) fprintf(stderr, "lut_index_lt8 mismatch at %a\n", ph); b++; } /* the PLL's clamp-free form, all of [0, 8) */
      }
      bad += b;
    });
  for (auto &x : th) x.join();
  /* AGC table index and int16 output truncation without FP64: every float in [0,1] / every float in (-4,4) */
  std::atomic<uint64_t> bad2(0);
  {
    std::vector<std::thread> t2;
    for (unsigned t = 0; t < nt; t++)
      t2.emplace_back([&, t]() {
        uint64_t b = 0;
        const uint32_t ONE = 0x3F800000u, FOUR = 0x40800000u;
        for (uint64_t u = (uint64_t)(ONE + 1) * t / nt; u < (uint64_t)(ONE + 1) * (t + 1) / nt; u++) {
          uint32_t bits = (uint32_t)u; float a; memcpy(&a, &bits, 4);
          if ((int)((double)a * 32767.0) != sdrk::RoleAgc::q15_index(a)) { if (b < 5) fprintf(stderr, "q15_index mismatch at %a\n", a); b++; }
        }
        for (uint64_t u = (uint64_t)FOUR * t / nt; u < (uint64_t)FOUR * (t + 1) / nt; u++) {
          for (int sgn = 0; sgn < 2; sgn++) {
            uint32_t bits = (uint32_t)u | (sgn ? 0x80000000u : 0u); float g; memcpy(&g, &bits, 4);
            int want = (int)(int16_t)(int)((double)g * 32767.0);
            if (want != sdrk::RoleOut::pcm(g)) { if (b < 5) fprintf(stderr, "pcm mismatch at %a: %d vs %d\n", g, want, sdrk::RoleOut::pcm(g)); b++; }
          }
        }
        bad2 += b;
      });
    for (auto &x : t2) x.join();
  }
  bad += bad2.load();
  /* large magnitudes and specials for pcm */
  {
    const float specials[] = {65535.0f, 65536.0f, 65536.5f, 65537.99f, 65538.0f, 70000.0f, 1e9f, 3e38f, -65535.9f, -65536.0f, -65537.0f, -65538.5f, -1e20f};
    for (float g : specials) {
      double d = (double)g * 32767.0;
      int i = (d >= 2147483648.0 || d <= -2147483649.0) ? (int)0x80000000 : (int)d;
      if ((int)(int16_t)i != sdrk::RoleOut::pcm(g)) { fprintf(stderr, "pcm special mismatch at %g\n", g); bad += 1; }
    }
  }
  /* float-only forms of (float)((double)x + PI/2), (float)((double)x +- PI) over their whole operand ranges, and the PLL wrap compares */
  {
    std::atomic<uint64_t> bad3(0);
    std::vector<std::thread> t3;
    for (unsigned t = 0; t < nt; t++)
      t3.emplace_back([&, t]() {
        uint64_t b = 0;
        auto walk = [&](float lim, int sign, int which) {
          uint32_t top; memcpy(&top, &lim, 4);
          for (uint64_t u = (uint64_t)(top + 1) * t / nt; u < (uint64_t)(top + 1) * (t + 1) / nt; u++) {
            uint32_t bits = (uint32_t)u | (sign ? 0x80000000u : 0u); float x; memcpy(&x, &bits, 4);
            float want, got;
            if (which == 0) {
              want = (float)((double)x + SDR_PI_D / 2.0); got = sdrk::add_half_pi(x);
              const float got2 = sdrk::add_half_pi_inrange(x); /* the PLL's Fast2Sum form */
              if (memcmp(&want, &got2, 4)) { if (b < 5) fprintf(stderr, "add_half_pi_inrange mismatch at %a\n", x); b++; }
              if ((got2 < 0.0f) != (x < -0x1.921fb4p+0f)) { if (b < 5) fprintf(stderr, "cosine wrap compare mismatch at %a\n", x); b++; }
              if (x >= -0x1.921fb6p+1f && x <= 0x1.921fb6p+1f && sdrk::lut_index_cos(x) != sdrk::lut_index_below_2pi(got2)) { if (b < 5) fprintf(stderr, "lut_index_cos mismatch at %a\n", x); b++; }
            }
            else if (which == 1) { want = (float)((double)x + SDR_PI_D); got = sdrk::add_pi(x); }
            else { want = (float)((double)x - SDR_PI_D); got = sdrk::sub_pi(x); }
            if (memcmp(&want, &got, 4)) { if (b < 5) fprintf(stderr, "add_dconst(%d) mismatch at %a\n", which, x); b++; }
            if (which == 0) {
              if (((double)x >= SDR_PI_D) != (x >= 0x1.921fb6p+1f) || ((double)x < -SDR_PI_D) != (x < -0x1.921fb4p+1f)) { if (b < 5) fprintf(stderr, "wrap compare mismatch at %a\n", x); b++; }
            }
          }
        };
        walk(6.2831860f, 0, 0); walk(3.1415930f, 1, 0);
        walk(0.79f, 0, 1); walk(0.79f, 1, 1); walk(0.79f, 0, 2); walk(0.79f, 1, 2);
        bad3 += b;
      });
    for (auto &x : t3) x.join();
    bad += bad3.load();
  }
  uint64_t badq = 0;
  for (int q = -32768; q <= 32767; q++) {
    double want = (double)(float)q / 32767.0, got = sdrk::RoleIn::q15_to_double(q);
    if (memcmp(&want, &got, 8) != 0) { if (badq < 5) fprintf(stderr, "q15 mismatch at %d: %a vs %a\n", q, want, got); badq++; }
  }
  bad += badq;
  printf("checked %u floats + 65536 int16 values, mismatches %llu\n", END, (unsigned long long)bad.load());
  return bad.load() ? 1 : 0;
}
