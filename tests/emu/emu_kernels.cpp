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

#include <algorithm>
#include <string>
#include <vector>
#include <stdio.h>

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

/* phase: 0 = load, 1 = step(t) part A, 3 = step(t) part B (stages that exchange data between lanes run in two
 * parts, the kernel separates them with __syncwarp()), 2 = save -- mirrors run_stage() of sdr_kernel.cu */
void dispatch(const Ctx &x, Warp &k, int w, int lane, int phase, uint32_t t) {
  if (phase == 3 && w != ST_IN && w != ST_OUT) return;
  x.k.set(*x.Y, phase == 0 ? 0u : (phase == 2 ? x.L->n_tiles : t)); /* the kernel's tile loop counts these */
  emu_async_owner() = w; /* asynchronous copies belong to the stage (warp) that issued them: its lanes wait for their own, then meet at a warp barrier */
  const bool ssb = x.Y->cls == CLS_SSB;
  if (w == ST_IN) { k.in.resize(32); RoleIn &r = k.in[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step_a(x, lane, t); else if (phase == 3) r.step_b(x, lane, t); else r.save(x, lane); }
  else if (w == ST_NB) { k.nbk.resize(32); RoleNb &r = k.nbk[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x, lane); }
  else if (w == ST_ENVL) { k.envl.resize(32); RoleEnvl &r = k.envl[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); }
  else if (w == ST_NBO) { k.nbo.resize(32); RoleNbo &r = k.nbo[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); }
  else if (w == ST_IFI || w == ST_IFQ || w == ST_AUD || (!ssb && (w == ST_IMGI || w == ST_IMGQ))) {
    k.bq.resize(32); RoleBiquad &r = k.bq[lane];
    const bool is_if = w == ST_IFI || w == ST_IFQ, is_aud = w == ST_AUD;
    if (phase == 0) r.load(x, lane, is_if ? 0 : (is_aud ? 1 : 2), is_if ? w - ST_IFI : (is_aud ? 0 : w - ST_IMGI));
    else if (phase == 1) r.step(x, lane, t);
    else r.save(x);
  } else if (w == ST_AGC) {
    k.agc.resize(32); RoleAgc &r = k.agc[lane];
    if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x);
  } else if (w == ST_OUT) {
    k.out.resize(32); RoleOut &r = k.out[lane];
    if (phase == 0) r.load(x, lane); else if (phase == 1) r.step_a(x, lane, t); else if (phase == 3) r.step_b(x, lane, t); else r.save(x, lane);
  } else if (ssb) {
    if (w == ST_NCO) {
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
    else { k.hil.resize(32); RoleHilbert &r = k.hil[lane]; const int sub = w - ST_HIL0;
      if (phase == 0) r.load(x, lane, sub); else if (phase == 1) r.step(x, g_hilbert, lane, sub, t); else r.save(x, lane, sub); }
  } else {
    if (w == ST_PLL) { k.pll.resize(32); RolePll &r = k.pll[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x); }
    else if (w == ST_NCO2) { k.nco2.resize(32); RoleNco2 &r = k.nco2[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x); }
    else { k.mag.resize(32); RoleMag &r = k.mag[lane]; if (phase == 0) r.load(x, lane); else if (phase == 1) r.step(x, lane, t); else r.save(x); }
  }
}

/* Scheduling policies (SDR_EMU_SCHED): the stages of a group only obey the hand-over rules of the launch's plan
 * (SdrLay::deps), so every order those rules allow must give the same bits.
 *   lockstep (default)  the lock-step schedule: at step s every stage with delay d runs tile s - d; the rules are asserted
 *                       (SDR_EMU_REVERSE=1 reverses the order of the stages inside a step)
 *   producers           always run the most upstream runnable stage: producers get as far ahead as the rules allow
 *   consumers           always run the most downstream runnable stage: producers run only when somebody needs them
 *   random:<seed>       a runnable stage picked at random
 * Shared memory is poisoned before every group, so a tile read before it was written or after it was overwritten shows up. */
static bool runnable(const SdrLay &Y, const std::vector<long long> &done, int s, uint32_t n) {
  if (!Y.active[s] || done[s] >= (long long)n) return false;
  const long long t = done[s];
  for (int i = 0; i < SDR_MAX_DEPS && Y.deps[s][i].stage >= 0; i++) {
    const SdrDep d = Y.deps[s][i];
    const long long u = d.kind ? (t | (long long)(Y.tpb - 1)) : t + d.k;
    if (u < 0) continue;
    for (int m = 0; m < SDR_STAGES; m++) /* a barrier completes when every stage of its group has arrived */
      if (Y.active[m] && Y.bar_of[m] == d.stage && done[m] <= u) return false;
  }
  return true;
}

int run_group(const SdrLaunch &L, const SdrGroup &G, int gidx) {
  const SdrLay &Y = L.lay;
  std::vector<unsigned char> smem((size_t)Y.smem_bytes, 0xFF);
  Ctx x; x.L = &L; x.Y = &L.lay; x.G = &G; x.smem = smem.data(); x.gidx = gidx; x.t0 = 0; x.prof = false;
  for (int i = 0; i < 257; i++) x.f(Y.o_sine)[i] = L.tabs->sine[i];
  for (int i = 0; i < 32; i++) reinterpret_cast<int *>(smem.data() + Y.o_cid)[i] = G.cid[i];
  for (int i = 0; i < SDR_LUT_SLOTS * SDR_AGC_LUT_STRIDE; i++) {
    const int id = G.lut_ids[i / SDR_AGC_LUT_STRIDE];
    if (id >= 0) x.f(Y.o_lut)[i] = L.agc_luts[(size_t)id * SDR_AGC_LUT_STRIDE + i % SDR_AGC_LUT_STRIDE];
  }
  const uint32_t n = L.n_tiles;
  std::vector<Warp> W(SDR_STAGES);
  std::vector<long long> done(SDR_STAGES, 0);
  std::vector<int> order; /* active stages, most upstream first */
  for (int d = 0; d <= Y.dmax; d++) for (int s = 0; s < SDR_STAGES; s++) if (Y.active[s] && Y.delay[s] == d) order.push_back(s);
  for (int s : order) for (int lane = 0; lane < 32; lane++) dispatch(x, W[s], s, lane, 0, 0);
  auto run_tile = [&](int s) {
    const uint32_t t = (uint32_t)done[s];
    for (int lane = 0; lane < 32; lane++) dispatch(x, W[s], s, lane, 1, t);
    for (int lane = 0; lane < 32; lane++) dispatch(x, W[s], s, lane, 3, t);
    done[s]++;
  };
  const char *pol = getenv("SDR_EMU_SCHED");
  std::string policy = pol && *pol ? pol : "lockstep";
  if (policy == "lockstep") {
    const char *r = getenv("SDR_EMU_REVERSE");
    const bool reverse = r && r[0] == '1';
    for (uint32_t s = 0; s < n + (uint32_t)Y.dmax; s++) {
      /* the warps of a step in either order; the stages one warp runs in a step (its program) always in program order */
      for (int wi = 0; wi < Y.n_warps; wi++) {
        const int w = reverse ? Y.n_warps - 1 - wi : wi;
        for (int i = 0; i < 4 && Y.prog[w][i] != 0xFF; i++) {
          const int st = Y.prog[w][i];
          if (!Y.active[st]) continue;
          const long long tau = (long long)s - Y.delay[st];
          if (tau < 0 || tau >= (long long)n) continue;
          if (done[st] != tau || !runnable(Y, done, st, n)) { fprintf(stderr, "[emu] lock-step schedule violates a hand-over rule: stage %d tile %lld\n", st, tau); return 1; }
          run_tile(st);
        }
      }
    }
  } else {
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    if (policy.rfind("random", 0) == 0 && policy.size() > 7) rng ^= strtoull(policy.c_str() + 7, nullptr, 10) * 0xBF58476D1CE4E5B9ull;
    for (;;) {
      std::vector<int> can;
      for (int st : order) if (runnable(Y, done, st, n)) can.push_back(st);
      if (can.empty()) break;
      int pick;
      if (policy == "producers") pick = can.front();
      else if (policy == "consumers") pick = can.back();
      else { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; pick = can[(size_t)(rng % can.size())]; }
      run_tile(pick);
    }
    for (int st : order) if (done[st] != (long long)n) { fprintf(stderr, "[emu] deadlock: stage %d stopped at tile %lld of %u\n", st, done[st], n); return 2; }
  }
  for (int s : order) for (int lane = 0; lane < 32; lane++) dispatch(x, W[s], s, lane, 2, 0);
  return 0;
}

/* the ALS + output post-pass (sdr_als_pass.cu): one warp in program order, lane by lane between the kernel's warp barriers */
int run_als_group(const SdrLaunch &L, const SdrGroup &G, int gidx) {
  const SdrLay &Y = L.lay;
  std::vector<unsigned char> smem((size_t)Y.smem_bytes, 0xFF);
  Ctx x; x.L = &L; x.Y = &L.lay; x.G = &G; x.smem = smem.data(); x.gidx = gidx; x.t0 = 0; x.prof = false;
  for (int i = 0; i < 32; i++) reinterpret_cast<int *>(smem.data() + Y.o_cid)[i] = G.cid[i];
  emu_async_owner() = ST_OUT;
  std::vector<RoleOut> out(32);
  x.k.reset();
  for (int lane = 0; lane < 32; lane++) out[lane].load(x, lane);
  const uint32_t n = L.n_tiles;
  for (int lane = 0; lane < 32; lane++) RoleAlsIn::request(x, lane, 0, 0);
  for (uint32_t t = 0; t < n; t++) {
    cp_async_wait_pending(0);
    if (t + 1 < n) for (int lane = 0; lane < 32; lane++) RoleAlsIn::request(x, lane, t + 1, Slots::next(x.k.c, Y.nc));
    for (int lane = 0; lane < 32; lane++) { if (Y.als_mirror) out[lane].step_a<true>(x, lane, t); else out[lane].step_a<false>(x, lane, t); }
    for (int lane = 0; lane < 32; lane++) out[lane].step_b(x, lane, t);
    x.k.advance(x);
  }
  for (int lane = 0; lane < 32; lane++) out[lane].save(x, lane);
  return 0;
}

}  // namespace

extern "C" {

int sdrk_setup_device(const float *hilbert64) { memcpy(g_hilbert, hilbert64, sizeof g_hilbert); return 0; }

int sdrk_launch_pipeline(const SdrLaunch *L, void *) {
  if (lay_check(&L->lay)) return 100 + lay_check(&L->lay);
  for (uint32_t g = 0; g < L->n_groups; g++) { int e = run_group(*L, L->groups[g], (int)g); if (e) return e; }
  return 0;
}

int sdrk_launch_als_pass(const SdrLaunch *L, void *) {
  if (L->lay.cls != CLS_ALS || !L->raw) return 99;
  for (uint32_t g = 0; g < L->n_groups; g++) { int e = run_als_group(*L, L->groups[g], (int)g); if (e) return e; }
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

static size_t emu_state_index(uint32_t w, uint32_t c, unsigned long long ch_stride) {
  if (w < W_NB_RING) return (size_t)w * ch_stride + c;
  const uint32_t r = w - W_NB_RING;
  return (size_t)W_NB_RING * ch_stride + ((size_t)(r >> 2) * ch_stride + c) * 4 + (r & 3u);
}

int sdrk_launch_gather(const float *state, unsigned long long ch_stride, const uint32_t *chan, uint32_t n, const uint32_t *words,
                       uint32_t n_words, float *out, void *) {
  for (uint32_t e = 0; e < n; e++)
    for (uint32_t k = 0; k < n_words; k++) out[(size_t)e * n_words + k] = state[emu_state_index(words ? words[k] : k, chan ? chan[e] : e, ch_stride)];
  return 0;
}

int sdrk_launch_scatter(float *state, unsigned long long ch_stride, const uint32_t *chan, uint32_t n, const float *in, void *) {
  for (uint32_t e = 0; e < n; e++)
    for (uint32_t k = 0; k < SDR_STATE_WORDS; k++) state[emu_state_index(k, chan[e], ch_stride)] = in[(size_t)e * SDR_STATE_WORDS + k];
  return 0;
}

int sdrk_occupancy(const SdrLaunch *) { return 0; }
}
