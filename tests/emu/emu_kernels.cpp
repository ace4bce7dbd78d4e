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
  std::vector<RoleIn> in; std::vector<RoleNb> nbk; std::vector<RoleBiquad> bq; std::vector<RoleNco> nco; std::vector<RoleHilbert> hil;
  std::vector<RoleAgc> agc; std::vector<RoleOut> out; std::vector<RolePll> pll; std::vector<RoleNco2> nco2; std::vector<RoleMag> mag;
  std::vector<RoleEnvl> envl; std::vector<RoleNbo> nbo;
};

int delay_of(int cls, int w) {
  static const int ssb[14] = {D_IN, D_NB, D_IF, D_IF, D_NCO, D_HIL, D_HIL, D_HIL, D_HIL, D_AUD, D_AGC, D_OUT, D_ENVL, D_NBO};
  static const int env[14] = {D_IN, D_NB, D_IF, D_IF, E_D_PLL, E_D_NCO2, E_D_IMG, E_D_IMG, E_D_MAG, E_D_AUD, E_D_AGC, E_D_OUT, D_ENVL, D_NBO};
  return cls == CLS_SSB ? ssb[w] : env[w];
}

/* phase: 0 = load, 1 = step(t) part A, 3 = step(t) part B (stages that exchange data between lanes run in two
 * parts, the kernel separates them with __syncwarp()), 2 = save -- mirrors run_group() of sdr_kernel.cu */
void dispatch(const Ctx &x, Warp &k, int w, int lane, int phase, uint32_t t) {
  if (phase == 3 && w != 0 && w != 11) return;
  const bool ssb = x.G->cls == CLS_SSB;
  const int oc = ssb ? (int)S_C : (int)E_C, oa = ssb ? (int)S_ALSC : (int)E_ALSC;
  if (w == 0) { k.in.resize(32); RoleIn &r = k.in[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step_a(x, lane, t); else if (phase == 3) r.step_b(x, lane, t); else r.save(x, lane); }
  else if (w == 1) { k.nbk.resize(32); RoleNb &r = k.nbk[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x, lane); }
  else if (w == 12) { k.envl.resize(32); RoleEnvl &r = k.envl[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); }
  else if (w == 13) { k.nbo.resize(32); RoleNbo &r = k.nbo[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); }
  else if (w == 2 || w == 3) {
    k.bq.resize(32); RoleBiquad &r = k.bq[lane]; const int rail = w - 2;
    if (phase == 0) r.load(x, lane, 0, rail);
    else if (phase == 1) r.step(x.tile(S_X, (t % NR) * 2 + rail), x.tile(S_Y, (t % NR) * 2 + rail), lane, true);
    else r.save(x, 0, rail);
  } else if (w == 9) {
    k.bq.resize(32); RoleBiquad &r = k.bq[lane];
    const int src = ssb ? (int)S_A : (int)E_A, dst = ssb ? (int)S_B : (int)E_B, nd = ssb ? (int)NA : (int)NB_RING;
    if (phase == 0) r.load(x, lane, 1, 0); else if (phase == 1) r.step(x.tile(src, t % nd), x.tile(dst, t % nd), lane, r.on); else r.save(x, 1, 0);
  } else if (w == 10) {
    k.agc.resize(32); RoleAgc &r = k.agc[lane];
    const int src = ssb ? (int)S_B : (int)E_B, ns = ssb ? (int)NA : (int)NB_RING;
    if (phase == 0) r.load(x, lane);
    else if (phase == 1) r.step(x.tile(src, t % ns), x.tile(oc, t % NC), lane, ssb ? 0.0f : x.f(E_CARR)[((t >> 2) & 7) * SDR_LANES + lane]);
    else r.save(x);
  } else if (w == 11) {
    k.out.resize(32); RoleOut &r = k.out[lane];
    if (phase == 0) r.load(x, lane, oc, oa); else if (phase == 1) r.step_a(x, lane, t, oc, oa); else if (phase == 3) r.step_b(x, lane, t); else r.save(x, lane, oc, oa);
  } else if (ssb) {
    if (w == 4) {
      k.nco.resize(32); RoleNco &r = k.nco[lane];
      if (phase == 0) {
        r.load(x, lane);
        if (lane == 31) { /* the warp vote of sdr_kernel.cu, on the 32 lane objects */
          int leader = -1;
          for (int l = 0; l < 32; l++) if (k.nco[l].cid >= 0) { leader = l; break; }
          bool uni = leader >= 0;
          for (int l = 0; l < 32 && uni; l++)
            if (k.nco[l].cid >= 0 && (f2u(k.nco[l].phase) != f2u(k.nco[leader].phase) || f2u(k.nco[l].inc) != f2u(k.nco[leader].inc))) uni = false;
          if (const char *e = getenv("SDR_EMU_NO_UNIFORM")) if (e[0] == '1') uni = false;
          for (int l = 0; l < 32; l++) { k.nco[l].uniform = uni; if (uni) { k.nco[l].phase = k.nco[leader].phase; k.nco[l].inc = k.nco[leader].inc; } }
        }
      } else if (phase == 1) {
        if (r.uniform) { if (lane == 0) for (int l = 0; l < 32; l++) k.nco[l].table_step(x, l); r.mix_step(x, lane, t); }
        else r.step(x, lane, t);
      } else r.save(x);
    }
    else { k.hil.resize(32); RoleHilbert &r = k.hil[lane]; const int sub = w - 5;
      if (phase == 0) r.load(x, lane, sub); else if (phase == 1) r.step(x, g_hilbert, lane, sub, t); else r.save(x, lane, sub); }
  } else {
    if (w == 4) { k.pll.resize(32); RolePll &r = k.pll[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x); }
    else if (w == 5) { k.nco2.resize(32); RoleNco2 &r = k.nco2[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x); }
    else if (w == 8) { k.mag.resize(32); RoleMag &r = k.mag[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x); }
    else { k.bq.resize(32); RoleBiquad &r = k.bq[lane]; const int rail = w - 6;
      if (phase == 0) r.load(x, lane, 2, rail);
      else if (phase == 1) r.step(x.tile(E_Z2, (t % NZ2) * 2 + rail), x.tile(E_V, (t % NZ2) * 2 + rail), lane, r.cid >= 0 && env_flag(x, lane, t) != 0);
      else r.save(x, 2, rail); }
  }
}

void run_group(const SdrLaunch &L, const SdrGroup &G, bool reverse) {
  std::vector<unsigned char> smem(SDR_SMEM_BYTES, 0xFF);
  Ctx x; x.L = &L; x.G = &G; x.smem = smem.data(); x.gidx = 0; x.t0 = 0; x.prof = false;
  for (int i = 0; i < 257; i++) x.f(S_SINE)[i] = L.tabs->sine[i];
  for (int i = 0; i < 32; i++) reinterpret_cast<int *>(smem.data() + S_CID)[i] = G.cid[i];
  for (int i = 0; i < SDR_LUT_SLOTS * SDR_AGC_LUT_STRIDE; i++) {
    const int id = G.lut_ids[i / SDR_AGC_LUT_STRIDE];
    if (id >= 0) x.f(S_LUT)[i] = L.agc_luts[(size_t)id * SDR_AGC_LUT_STRIDE + i % SDR_AGC_LUT_STRIDE];
  }
  const int cls = G.cls;
  const uint32_t n = L.n_tiles;
  std::vector<Warp> W(SDR_WARPS);
  for (int w = 0; w < SDR_WARPS; w++) for (int lane = 0; lane < 32; lane++) dispatch(x, W[w], w, lane, 0, 0);
  const int dmax = cls == CLS_SSB ? D_SSB_MAX : D_ENV_MAX;
  for (uint32_t s = 0; s < n + (uint32_t)dmax; s++) {
    for (int wi = 0; wi < SDR_WARPS; wi++) {
      int w = reverse ? SDR_WARPS - 1 - wi : wi;
      long long tau = (long long)s - delay_of(cls, w);
      if (tau < 0 || tau >= (long long)n) continue;
      for (int lane = 0; lane < 32; lane++) dispatch(x, W[w], w, lane, 1, (uint32_t)tau);
      for (int lane = 0; lane < 32; lane++) dispatch(x, W[w], w, lane, 3, (uint32_t)tau);
    }
  }
  for (int w = 0; w < SDR_WARPS; w++) for (int lane = 0; lane < 32; lane++) dispatch(x, W[w], w, lane, 2, 0);
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
      if ((m & SDRK_R_NB) && w >= W_NB_MASK && w < W_NB_RING) z = true;
      if (z) state[(size_t)w * ch_stride + c] = 0.0f;
    }
    if (m & SDRK_R_NB) {
      float4 *ring = reinterpret_cast<float4 *>(state + (size_t)W_NB_RING * ch_stride);
      for (uint32_t f = 0; f < 288; f++) { float4 z4; z4.x = z4.y = z4.z = z4.w = 0.0f; ring[(size_t)f * ch_stride + c] = z4; }
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
