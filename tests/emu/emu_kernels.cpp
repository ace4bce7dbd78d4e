/* tests/emu/emu_kernels.cpp -- TEST SCAFFOLDING, never shipped.
 *
 * Host stand-in for audiosdr_b200/csrc/sdr_kernel.cu: the same role bodies (sdr_pipeline.cuh, compiled
 * by g++) are run lane by lane, warp by warp, step by step, with the same barrier structure, on host
 * memory.  It lets the pipeline LOGIC (delays, ring slots, state carry, reset replay, grouping) be
 * checked against the oracle in the CPU-only test tier, where no GPU exists.  The GPU tier then checks
 * the real kernels.  "Shared memory" is poisoned with 0xFF before every group so that a read of a
 * tile that no stage has written shows up as NaN / a bad mask code instead of a lucky zero, and the
 * order in which the warps of one step run can be reversed (SDR_EMU_REVERSE=1) to expose a tile that
 * is read and written in the same step.
 */
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../audiosdr_b200/csrc/sdr_kernel.h"
#include "../../audiosdr_b200/csrc/sdr_pipeline.cuh"

using namespace sdrk;

static float g_hilbert[64];

namespace {

struct Warp {
  int id;
  std::vector<RoleIn> in; std::vector<RoleBiquad> bq; std::vector<RoleNco> nco; std::vector<RoleHilbert> hil;
  std::vector<RoleAgc> agc; std::vector<RoleOut> out; std::vector<RolePll> pll; std::vector<RoleNco2> nco2; std::vector<RoleMag> mag;
};

int delay_of(int cls, int w) {
  static const int ssb[11] = {D_IN, D_IF, D_IF, D_NCO, D_HIL, D_HIL, D_HIL, D_HIL, D_AUD, D_AGC, D_OUT};
  static const int env[11] = {D_IN, D_IF, D_IF, E_D_PLL, E_D_NCO2, E_D_IMG, E_D_IMG, E_D_MAG, E_D_AUD, E_D_AGC, E_D_OUT};
  return cls == CLS_SSB ? ssb[w] : env[w];
}

void run_group(const SdrLaunch &L, const SdrGroup &G, bool reverse) {
  std::vector<unsigned char> smem(SDR_SMEM_BYTES, 0xFF);
  Ctx x; x.L = &L; x.G = &G; x.smem = smem.data();
  for (int i = 0; i < 257; i++) x.f(S_SINE)[i] = L.tabs->sine[i];
  const int cls = G.cls;
  const uint32_t n = L.n_tiles;
  std::vector<Warp> W(SDR_WARPS);
  /* load phase */
  for (int w = 0; w < SDR_WARPS; w++) {
    Warp &k = W[w]; k.id = w;
    for (int lane = 0; lane < 32; lane++) {
      if (w == 0) { k.in.resize(32); k.in[lane].load(x, lane); }
      else if (w == 1 || w == 2) { k.bq.resize(32); k.bq[lane].load(x, lane, 0, w - 1); }
      else if (w == 8) { k.bq.resize(32); k.bq[lane].load(x, lane, 1, 0); }
      else if (w == 9) { k.agc.resize(32); k.agc[lane].load(x, lane); }
      else if (w == 10) { k.out.resize(32); k.out[lane].load(x, lane, cls == CLS_SSB ? S_C : E_C, cls == CLS_SSB ? S_ALSC : E_ALSC); }
      else if (cls == CLS_SSB) {
        if (w == 3) { k.nco.resize(32); k.nco[lane].load(x, lane); }
        else { k.hil.resize(32); k.hil[lane].load(x, lane, w - 4); }
      } else {
        if (w == 3) { k.pll.resize(32); k.pll[lane].load(x, lane); }
        else if (w == 4) { k.nco2.resize(32); k.nco2[lane].load(x, lane); }
        else if (w == 5 || w == 6) { k.bq.resize(32); k.bq[lane].load(x, lane, 2, w - 5); }
        else { k.mag.resize(32); k.mag[lane].load(x, lane); }
      }
    }
  }
  const int dmax = cls == CLS_SSB ? D_SSB_MAX : D_ENV_MAX;
  for (uint32_t s = 0; s < n + (uint32_t)dmax; s++) {
    for (int wi = 0; wi < SDR_WARPS; wi++) {
      int w = reverse ? SDR_WARPS - 1 - wi : wi;
      long long tau = (long long)s - delay_of(cls, w);
      if (tau < 0 || tau >= (long long)n) continue;
      uint32_t t = (uint32_t)tau;
      Warp &k = W[w];
      for (int lane = 0; lane < 32; lane++) {
        if (w == 0) k.in[lane].step(x, lane, t);
        else if (w == 1 || w == 2) k.bq[lane].step(x.tile(S_X, (t & 1) * 2 + (w - 1)), x.tile(S_Y, (t & 1) * 2 + (w - 1)), lane, true);
        else if (cls == CLS_SSB) {
          if (w == 3) k.nco[lane].step(x, lane, t);
          else if (w <= 7) k.hil[lane].step(x, g_hilbert, lane, w - 4, t);
          else if (w == 8) k.bq[lane].step(x.tile(S_A, t & 1), x.tile(S_B, t & 1), lane, k.bq[lane].on);
          else if (w == 9) k.agc[lane].step(x.tile(S_B, t & 1), x.tile(S_C, t % NC), lane, 0.0f);
          else k.out[lane].step(x, lane, t, S_C, S_ALSC);
        } else {
          if (w == 3) k.pll[lane].step(x, lane, t);
          else if (w == 4) k.nco2[lane].step(x, lane, t);
          else if (w == 5 || w == 6)
            k.bq[lane].step(x.tile(E_Z2, (t & 1) * 2 + (w - 5)), x.tile(E_V, (t & 1) * 2 + (w - 5)), lane,
                            k.bq[lane].cid >= 0 && env_flag(x, lane, t) != 0);
          else if (w == 7) k.mag[lane].step(x, lane, t);
          else if (w == 8) k.bq[lane].step(x.tile(E_A, t & 1), x.tile(E_B, t % NB_RING), lane, k.bq[lane].on);
          else if (w == 9) k.agc[lane].step(x.tile(E_B, t % NB_RING), x.tile(E_C, t % NC), lane, x.f(E_CARR)[((t >> 2) & 7) * SDR_LANES + lane]);
          else k.out[lane].step(x, lane, t, E_C, E_ALSC);
        }
      }
    }
  }
  /* save phase */
  for (int w = 0; w < SDR_WARPS; w++) {
    Warp &k = W[w];
    for (int lane = 0; lane < 32; lane++) {
      if (w == 0) k.in[lane].save(x, lane);
      else if (w == 1 || w == 2) k.bq[lane].save(x, 0, w - 1);
      else if (w == 8) k.bq[lane].save(x, 1, 0);
      else if (w == 9) k.agc[lane].save(x);
      else if (w == 10) k.out[lane].save(x, lane, cls == CLS_SSB ? S_C : E_C, cls == CLS_SSB ? S_ALSC : E_ALSC);
      else if (cls == CLS_SSB) {
        if (w == 3) k.nco[lane].save(x);
        else k.hil[lane].save(x, lane, w - 4);
      } else {
        if (w == 3) k.pll[lane].save(x);
        else if (w == 4) k.nco2[lane].save(x);
        else if (w == 5 || w == 6) k.bq[lane].save(x, 2, w - 5);
        else k.mag[lane].save(x);
      }
    }
  }
}

}  // namespace

extern "C" {

int sdrk_setup_device(const float *hilbert64) { memcpy(g_hilbert, hilbert64, sizeof g_hilbert); return 0; }

int sdrk_launch_pipeline(const SdrLaunch *L, void *) {
  const char *r = getenv("SDR_EMU_REVERSE");
  bool reverse = r && r[0] == '1';
  for (uint32_t g = 0; g < L->n_groups; g++) run_group(*L, L->groups[g], reverse);
  return 0;
}

int sdrk_launch_reset(float *state, unsigned long long ch_stride, const uint32_t *chan, const uint32_t *mask, uint32_t n, void *) {
  for (uint32_t e = 0; e < n; e++) {
    uint32_t c = chan[e], m = mask[e];
    for (uint32_t w = 0; w < SDR_STATE_WORDS; w++) {
      bool z = false;
      if ((m & SDRK_R_IF) && w < W_IF_Q + 16) z = true;
      if ((m & SDRK_R_IMG) && w >= W_IMG_I && w < W_IMG_Q + 16) z = true;
      if ((m & SDRK_R_AUD) && w >= W_AUD && w < W_AUD + 16) z = true;
      if ((m & SDRK_R_ALS) && w >= W_ALS_C && w < W_ALS_H + 128) z = true;
      if ((m & SDRK_R_NB) && w >= W_NB_MASK && w < W_NB_RING + 768) z = true;
      if (z) state[(size_t)w * ch_stride + c] = 0.0f;
    }
  }
  return 0;
}

int sdrk_launch_fill_word(float *state, unsigned long long ch_stride, uint32_t w, float v, uint32_t n_ch, void *) {
  for (uint32_t c = 0; c < n_ch; c++) state[(size_t)w * ch_stride + c] = v;
  return 0;
}

int sdrk_launch_gather(const float *state, unsigned long long ch_stride, const uint32_t *chan, uint32_t n, const uint32_t *words,
                       uint32_t n_words, float *out, void *) {
  for (uint32_t e = 0; e < n; e++)
    for (uint32_t k = 0; k < n_words; k++) out[(size_t)e * n_words + k] = state[(size_t)words[k] * ch_stride + (chan ? chan[e] : e)];
  return 0;
}
}
