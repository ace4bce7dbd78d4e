/* sdr_lay.h -- shared-memory plan and hand-over rules of one pipeline launch.
 *
 * A launch covers the groups of ONE bucket: a pipeline class (SSB: NCO + Hilbert; ENV: PLL + envelope) with one set of
 * optional stages (blanker trio; ALS history).  For that bucket lay_build() decides
 *   - the tile length T (samples a stage works on at a time: 32, 16 or 8; 128 / T tiles make one reference block),
 *   - which stages exist and which warp runs which stage,
 *   - where every ring and table lives in the CTA's shared memory and how many tiles each ring holds,
 *   - for every stage, which tiles of which other stages it has to wait for (SdrDep).
 * The stages of a group are NOT stepped in lock step: stage X starts tile t as soon as every stage it depends on has
 * finished the tiles named by its rules -- producers for data (read-after-write), consumers for ring slots
 * (write-after-read) -- signalled through one mbarrier per (stage, tile mod SDR_BAR_W).  The lock-step schedule "stage
 * with delay d works on tile s - d at step s" is one valid order of that partial order (lay_check() proves it for the
 * rules built here: every rule points to an earlier step), so the rules cannot deadlock, and ring depths beyond the
 * minimum only add slack between neighbours.
 *
 * Plain C-style header: compiled into the product (sdr_host.cpp, sdr_kernel.cu) and into the host emulation of the
 * test suite (tests/emu), which runs the same rules with adversarial schedules.
 */
#ifndef SDR_LAY_H
#define SDR_LAY_H
#include <stdint.h>
#include <string.h>

#include "sdr_types.h"

/* stage ids (= the `stage` argument of run_stage; a launch's warps map onto a subset of them) */
enum {
  ST_IN = 0, ST_NB = 1, ST_IFI = 2, ST_IFQ = 3,
  ST_NCO = 4, ST_HIL0 = 5, /* .. ST_HIL0 + 3 */
  ST_PLL = 4, ST_NCO2 = 5, ST_IMGI = 6, ST_IMGQ = 7, ST_MAG = 8,
  ST_AUD = 9, ST_AGC = 10, ST_OUT = 11, ST_ENVL = 12, ST_NBO = 13
};

/* bucket features.  LF_SAM: an ENV bucket whose groups hold SAM channels only (no AM lane): the AGC never needs the
 * block's carrier level (C:408-409 is AM-mode only), and the envelope path runs only for blocks the PLL ends unlocked
 * (C:130-132), i.e. rarely -- so on short tiles its four stages share one warp, and input and output another: 7 warps
 * instead of 11, three groups per SM. */
enum { LF_NB = 1u, LF_ALS = 2u, LF_SAM = 4u };

/* The plan of every 32-sample-tile launch is the same (whatever optional stages the bucket has): offsets and ring depths
 * are compile-time constants of the 32-sample kernel, and a group with all stages fills the SM's shared memory. */
enum {
#ifdef SDR_HANDOVER /* (experiment build: the hand-over barriers need room, taken from the Hilbert Q ring) */
  LAY32_HQ_TILES = 12, LAY32_BAR_BYTES = SDR_STAGES * SDR_BAR_W * 8,
#else
  LAY32_HQ_TILES = 16, LAY32_BAR_BYTES = 0, /* 16 tiles = 512 samples: ring positions wrap with a mask */
#endif
  LAY32_NR = 5, LAY32_NI = 6, LAY32_NA_SSB = 3, LAY32_NA_ENV = 5, LAY32_NC = 6, LAY32_NZ = 5, LAY32_NZ2 = 3,
  LAY32_TILE = 32 * SDR_LANES * 4,
  LAY32_SINE = 0, LAY32_LUT = 1152, LAY32_NCOT = 3328, LAY32_CID = 3584, LAY32_BAR = 3712,
  LAY32_NBS = LAY32_BAR + LAY32_BAR_BYTES,
  LAY32_INS = LAY32_NBS + 32 * SDR_LANES * 16,
  LAY32_OUTS = LAY32_INS + 2 * SDR_LANES * 36 * 4,
  LAY32_R = LAY32_OUTS + SDR_LANES * 36 * 4,
  LAY32_CLASS = LAY32_R + LAY32_NR * 2 * LAY32_TILE,
  /* SSB */
  LAY32_HQ = LAY32_CLASS,
  LAY32_HI = LAY32_HQ + (LAY32_HQ_TILES * 16 + SDR_HQ_MIRROR) * SDR_LANES * 8,
  LAY32_SA = LAY32_HI + LAY32_NI * LAY32_TILE,
  LAY32_SC = LAY32_SA + LAY32_NA_SSB * LAY32_TILE,
  LAY32_SMASK = LAY32_SC + LAY32_NC * LAY32_TILE,
  LAY32_SALSC = LAY32_SMASK + 3 * 128 * SDR_LANES,
  LAY32_SSB_END = LAY32_SALSC + 128 * SDR_LANES * 4,
  /* ENV */
  LAY32_Z = LAY32_CLASS,
  LAY32_Z2 = LAY32_Z + LAY32_NZ * 2 * LAY32_TILE,
  LAY32_EA = LAY32_Z2 + LAY32_NZ2 * 2 * LAY32_TILE,
  LAY32_EC = LAY32_EA + LAY32_NA_ENV * LAY32_TILE,
  LAY32_EMASK = LAY32_EC + LAY32_NC * LAY32_TILE,
  LAY32_EALSC = LAY32_EMASK + 3 * 128 * SDR_LANES,
  LAY32_FLAGS = LAY32_EALSC + 128 * SDR_LANES * 4,
  LAY32_CARR = LAY32_FLAGS + 8 * SDR_LANES * 4,
  LAY32_ENV_END = LAY32_CARR + 8 * SDR_LANES * 4
};

static inline int lay_align(int v, int a) { return (v + a - 1) / a * a; }

#ifdef SDR_EMU
#include <stdlib.h>
/* test scaffold only (tests/emu): SDR_EMU_DROP_RULE=n leaves out the n-th rule of every plan, to show that the
 * adversarial schedules of the emulation notice a missing rule */
static int lay_emu_rule_counter = 0;
static int lay_emu_dropped[4] = {-1, -1, -1, -1}; /* stage, barrier owner, kind, k of the rule left out */
#endif

static inline void lay_dep(SdrLay *L, int stage, int on, int kind, int k) {
  if (!L->active[on] || !L->active[stage]) return;
  on = L->bar_of[on];
  for (int i = 0; i < SDR_MAX_DEPS; i++) /* the members of a barrier group add the same rule once */
    if (L->deps[stage][i].stage == on && L->deps[stage][i].kind == kind && L->deps[stage][i].k == k) return;
#ifdef SDR_EMU
  if (lay_emu_dropped[0] == stage && lay_emu_dropped[1] == on && lay_emu_dropped[2] == kind && lay_emu_dropped[3] == k) return;
  { const char *e = getenv("SDR_EMU_DROP_RULE");
    if (e && *e && atoi(e) == lay_emu_rule_counter++) { lay_emu_dropped[0] = stage; lay_emu_dropped[1] = on; lay_emu_dropped[2] = kind; lay_emu_dropped[3] = k; return; } }
#endif
  for (int i = 0; i < SDR_MAX_DEPS; i++)
    if (L->deps[stage][i].stage < 0) { L->deps[stage][i].stage = (int8_t)on; L->deps[stage][i].kind = (int8_t)kind; L->deps[stage][i].k = (int16_t)k; return; }
  L->error = 1; /* table too small */
}

/* The rule set as ring depths stand in *L (called by lay_build after the depths are final). */
static inline void lay_rules(SdrLay *L) {
#ifdef SDR_EMU
  lay_emu_rule_counter = 0; lay_emu_dropped[0] = -1;
#endif
  for (int s = 0; s < SDR_STAGES; s++) for (int i = 0; i < SDR_MAX_DEPS; i++) { L->deps[s][i].stage = -1; L->deps[s][i].kind = 0; L->deps[s][i].k = 0; }
  const int nb = (L->feat & LF_NB) != 0, tpb = L->tpb;
  const int x0 = L->cls == CLS_SSB ? ST_NCO : ST_PLL; /* the last stage that touches an input-ring slot */
  /* input ring R (in place): IN -> [ENVL reads, NB-out overwrites] -> IF-I, IF-Q -> NCO / PLL */
  lay_dep(L, ST_IN, x0, 0, -L->nr);
  if (nb) {
    /* blanker, C:606-650.  HBM ring planes: IN writes I/Q of block B into slot B % 3, whose previous content (block B - 3)
     * NB-out fetched one block (tpb tiles) ago; ENVL does the same for the envelope plane, which the scan fetches for the last
     * time with the first tile of the previous block.  Mask slots in shared memory: NB-out reads the final mask of tile t once the scan
     * of tile t is over; the scan recycles the oldest mask slot two tiles after NB-out has left it (see RoleNb). */
    lay_dep(L, ST_IN, ST_NBO, 0, -tpb);
    lay_dep(L, ST_ENVL, ST_IN, 0, 0);    /* (and through it, IN(t) <- NB-out(t - tpb) <- NB(t - tpb): the scan has fetched the envelopes ENVL(t) replaces) */
    lay_dep(L, ST_NB, ST_NBO, 0, -2);    /* (also covers the envelopes the scan requests at the end of tile t for tile t + 1: ENVL(t - 2) wrote
                                            the latest of them, and NB-out(t - 2) has waited for ENVL(t - 2)) */
    lay_dep(L, ST_NBO, ST_NB, 0, -1);    /* the mask words NB-out(t) reads were last touched by the scan of tile t - 1 (windows and edges reach back into the oldest block, see RoleNb) */
    lay_dep(L, ST_NBO, ST_ENVL, 0, 0);   /* overwrites the slot ENVL reads */
    lay_dep(L, ST_IFI, ST_NBO, 0, 0);
    lay_dep(L, ST_IFQ, ST_NBO, 0, 0);
  } else {
    lay_dep(L, ST_IFI, ST_IN, 0, 0);
    lay_dep(L, ST_IFQ, ST_IN, 0, 0);
  }
  lay_dep(L, x0, ST_IFI, 0, 0);
  lay_dep(L, x0, ST_IFQ, 0, 0);
  const int back_c = (L->feat & LF_ALS) ? tpb : 0; /* tiles of AGC output the ALS stage reaches back (C:336, M + delay <= 129) */
  if (L->cls == CLS_SSB) {
    /* Hilbert rings: tile u of the Q ring is read by HIL(u .. u + back_q), tile u of the I ring by HIL(u + tpb) */
    const int back_q = (255 + L->T - 1) / L->T;
    int lead = L->hq_tiles - back_q;
    if (L->ni - tpb < lead) lead = L->ni - tpb;
    for (int h = 0; h < L->n_hil; h++) {
      lay_dep(L, ST_NCO, ST_HIL0 + h, 0, -lead);
      lay_dep(L, ST_HIL0 + h, ST_NCO, 0, 0);
      lay_dep(L, ST_HIL0 + h, ST_AGC, 0, -L->na);
      lay_dep(L, ST_AUD, ST_HIL0 + h, 0, 0);
    }
    lay_dep(L, ST_AGC, ST_AUD, 0, 0);
  } else {
    /* The envelope path of a block runs iff the PLL is unlocked after the block's LAST sample (C:130-132): NCO2 waits
     * for the block's last PLL tile; AM-mode AGC uses the carrier level after the block's envelope loop (C:408-409). */
    lay_dep(L, ST_PLL, ST_NCO2, 0, -L->nz);
    lay_dep(L, ST_NCO2, ST_PLL, 1, 0);
    lay_dep(L, ST_NCO2, ST_MAG, 0, -L->nz2);
    lay_dep(L, ST_IMGI, ST_NCO2, 0, 0);
    lay_dep(L, ST_IMGQ, ST_NCO2, 0, 0);
    lay_dep(L, ST_MAG, ST_IMGI, 0, 0);
    lay_dep(L, ST_MAG, ST_IMGQ, 0, 0);
    lay_dep(L, ST_MAG, ST_AGC, 0, -L->na);
    lay_dep(L, ST_AUD, ST_MAG, 0, 0);
    lay_dep(L, ST_AGC, ST_AUD, 0, 0);
    if (!(L->feat & LF_SAM)) lay_dep(L, ST_AGC, ST_MAG, 1, 0);
  }
  lay_dep(L, ST_AGC, ST_OUT, 0, -(L->nc - back_c));
  lay_dep(L, ST_OUT, ST_AGC, 0, 0);
}

/* Every rule must point to an earlier step of the lock-step schedule (deadlock freedom, see the header comment).
 * Barrier phases: stage P signals tile u on barrier (P, u mod SDR_BAR_W), and a waiter tells "tile u done" from "not yet" by
 * the barrier's phase parity, which is only unambiguous while P cannot finish tile u + SDR_BAR_W before the waiter has seen
 * tile u.  A consumer is never ahead of its producer and a producer is never more than the ring depth ahead of its
 * consumer, so it suffices that every ring depth and every backward distance stays below SDR_BAR_W - 2. */
static inline int lay_check(const SdrLay *L) {
  if (L->error) return 1;
  for (int s = 0; s < SDR_STAGES; s++) {
    if (!L->active[s]) continue;
    for (int i = 0; i < SDR_MAX_DEPS && L->deps[s][i].stage >= 0; i++) {
      const SdrDep d = L->deps[s][i];
      const int k = d.kind ? L->tpb - 1 : d.k; /* worst case of (t | (tpb - 1)) - t */
      if (!(k + L->delay[d.stage] < L->delay[s])) {
        /* the same step is in order only inside one warp's program, when the stage waited for comes first */
        int ok = 0;
        for (int w = 0; w < L->n_warps && !ok; w++) {
          int ps = -1, pd = -1;
          for (int i = 0; i < 4 && L->prog[w][i] != 0xFF; i++) { if (L->prog[w][i] == s) ps = i; if (L->bar_of[L->prog[w][i]] == d.stage) pd = i > pd ? i : pd; }
          if (ps >= 0 && pd >= 0 && pd < ps && k + L->delay[d.stage] == L->delay[s]) ok = 1;
        }
        if (!ok) return 2;
      }
      if (d.k < -(SDR_BAR_W - 2)) return 3;
    }
  }
  const int lim = SDR_BAR_W - 2;
  if (L->nr > lim || L->na > lim || L->nc > lim || L->nz > lim || L->nz2 > lim || L->ni > lim || L->tpb + 2 > lim) return 3;
  if (L->smem_bytes > 232448) return 4;
  if ((L->o_ins | L->o_outs | (L->ins_row * 4) | L->o_lut) & 15) return 5; /* bulk copies move 16-byte aligned rows */
  return 0;
}

/* slack: extra ring slots beyond the lock-step minimum, each ring at most `max_slack`, as long as `budget` bytes allow */
static inline int lay_build_ex(SdrLay *L, int cls, uint32_t feat, int T, int budget, int max_slack, int in_depth_override, const uint8_t *merged_order) {
  memset(L, 0, sizeof *L);
  if (T != 32 && T != 16 && T != 8) return 1;
  if ((feat & (LF_NB | LF_ALS)) && T != 32) return 1; /* the blanker scan and the ALS passes are written for 32-sample tiles */
  L->cls = cls; L->feat = feat; L->T = T; L->tpb = 128 / T; L->tile_f = T * SDR_LANES;
  for (int v = L->tpb; v > 1; v >>= 1) L->tpb_sh++;
  const int tpb = L->tpb, nb = (feat & LF_NB) != 0, als = (feat & LF_ALS) != 0, tile_b = L->tile_f * 4;
  L->n_hil = cls == CLS_SSB ? T / 8 : 0;
  /* stages and their lock-step delays */
  int8_t *d = L->delay;
  L->active[ST_IN] = 1; d[ST_IN] = 0;
  if (nb) { L->active[ST_ENVL] = L->active[ST_NB] = L->active[ST_NBO] = 1; d[ST_ENVL] = 1; d[ST_NB] = 1; d[ST_NBO] = 2; }
  L->active[ST_IFI] = L->active[ST_IFQ] = 1; d[ST_IFI] = d[ST_IFQ] = (int8_t)(nb ? 3 : 1);
  if (cls == CLS_SSB) {
    L->active[ST_NCO] = 1; d[ST_NCO] = (int8_t)(d[ST_IFI] + 1);
    for (int h = 0; h < L->n_hil; h++) { L->active[ST_HIL0 + h] = 1; d[ST_HIL0 + h] = (int8_t)(d[ST_NCO] + 1); }
    d[ST_AUD] = (int8_t)(d[ST_NCO] + 2); d[ST_AGC] = (int8_t)(d[ST_AUD] + 1);
  } else {
    L->active[ST_PLL] = L->active[ST_NCO2] = L->active[ST_IMGI] = L->active[ST_IMGQ] = L->active[ST_MAG] = 1;
    d[ST_PLL] = (int8_t)(d[ST_IFI] + 1); d[ST_NCO2] = (int8_t)(d[ST_PLL] + tpb); d[ST_IMGI] = d[ST_IMGQ] = (int8_t)(d[ST_NCO2] + 1);
    d[ST_MAG] = (int8_t)(d[ST_NCO2] + 2); d[ST_AUD] = (int8_t)(d[ST_MAG] + 1); d[ST_AGC] = (int8_t)(d[ST_MAG] + tpb);
  }
  /* SAM-only ENV bucket on short tiles: the four envelope stages run one after the other in one warp at the same step (their
   * work ring needs one slot), and the AGC follows the audio filter directly */
  const int merged = cls == CLS_ENV && (feat & LF_SAM) && !(feat & (LF_NB | LF_ALS)) && T != 32;
  if (merged) { d[ST_IMGI] = d[ST_IMGQ] = d[ST_MAG] = d[ST_NCO2]; d[ST_AUD] = (int8_t)(d[ST_NCO2] + 1); d[ST_AGC] = (int8_t)(d[ST_AUD] + 1); }
  L->active[ST_AUD] = L->active[ST_AGC] = L->active[ST_OUT] = 1; d[ST_OUT] = (int8_t)(d[ST_AGC] + 1);
  for (int s = 0; s < SDR_STAGES; s++) if (L->active[s] && d[s] > L->dmax) L->dmax = d[s];
  /* barrier groups */
  for (int s = 0; s < 16; s++) { L->bar_of[s] = (uint8_t)s; L->bar_count[s] = 1; }
  L->bar_of[ST_IFQ] = ST_IFI; L->bar_count[ST_IFI] = 2;
  if (cls == CLS_SSB) { for (int h = 1; h < L->n_hil; h++) L->bar_of[ST_HIL0 + h] = ST_HIL0; L->bar_count[ST_HIL0] = (uint8_t)L->n_hil; }
  else { L->bar_of[ST_IMGQ] = ST_IMGI; L->bar_count[ST_IMGI] = 2; }
  if (T == 32) {
    /* the fixed plan (LAY32_*): lock-step delays as if every optional stage existed, so that the ring depths are the same
     * for every bucket */
    d[ST_IN] = 0; d[ST_ENVL] = 1; d[ST_NB] = 1; d[ST_NBO] = 2; d[ST_IFI] = d[ST_IFQ] = 3;
    if (cls == CLS_SSB) { d[ST_NCO] = 4; for (int h = 0; h < 4; h++) d[ST_HIL0 + h] = 5; d[ST_AUD] = 6; d[ST_AGC] = 7; d[ST_OUT] = 8; }
    else { d[ST_PLL] = 4; d[ST_NCO2] = 8; d[ST_IMGI] = d[ST_IMGQ] = 9; d[ST_MAG] = 10; d[ST_AUD] = 11; d[ST_AGC] = 14; d[ST_OUT] = 15; }
    L->dmax = d[ST_OUT];
    L->nr = LAY32_NR; L->nc = LAY32_NC; L->ins_row = 36; L->in_depth = 1;
    L->o_sine = LAY32_SINE; L->o_lut = LAY32_LUT; L->o_ncot = LAY32_NCOT; L->o_cid = LAY32_CID; L->o_bar = LAY32_BAR; L->o_nbs = LAY32_NBS;
    L->o_ins = LAY32_INS; L->o_outs = LAY32_OUTS; L->o_r = LAY32_R;
    if (cls == CLS_SSB) {
      L->ni = LAY32_NI; L->hq_tiles = LAY32_HQ_TILES; L->hq_rows = LAY32_HQ_TILES * 16; L->na = LAY32_NA_SSB;
      L->o_hq = LAY32_HQ; L->o_hi = LAY32_HI; L->o_a = LAY32_SA; L->o_c = LAY32_SC; L->o_mask = LAY32_SMASK; L->o_alsc = LAY32_SALSC;
      L->smem_bytes = als ? LAY32_SSB_END : (nb ? LAY32_SALSC : LAY32_SMASK);
    } else {
      L->nz = LAY32_NZ; L->nz2 = LAY32_NZ2; L->na = LAY32_NA_ENV;
      L->o_z = LAY32_Z; L->o_z2 = LAY32_Z2; L->o_a = LAY32_EA; L->o_c = LAY32_EC; L->o_mask = LAY32_EMASK; L->o_alsc = LAY32_EALSC;
      L->o_flags = LAY32_FLAGS; L->o_carr = LAY32_CARR;
      L->smem_bytes = LAY32_ENV_END;
    }
    (void)budget; (void)max_slack;
  } else {
  /* minimum ring depths: a tile's slot lives from the writer's step to the last reader's step */
  const int x0 = cls == CLS_SSB ? ST_NCO : ST_PLL;
  const int back_q = (255 + T - 1) / T, back_c = als ? tpb : 0;
  L->nr = d[x0] + 1;
  L->nc = back_c + 2;
  if (cls == CLS_SSB) { L->hq_tiles = back_q + 2; L->ni = tpb + 2; L->na = 3; }
  else { L->nz = tpb + 1; L->nz2 = merged ? 1 : 3; L->na = merged ? 3 : tpb + 1; }
  /* fixed part */
  int o = 0;
  L->o_sine = o; o += 1152;
  L->o_lut = o; o += lay_align(SDR_LUT_SLOTS * SDR_AGC_LUT_STRIDE * 4, 128);
  if (cls == CLS_SSB) { L->o_ncot = o; o += 256; } /* the shared NCO table of an SSB group */
  L->o_cid = o; o += 128;
  L->o_bar = o; o += LAY32_BAR_BYTES; /* hand-over barriers: only the -DSDR_HANDOVER build has them */
  if (cls == CLS_ENV) { L->o_flags = o; o += 8 * SDR_LANES * 4; L->o_carr = o; o += 8 * SDR_LANES * 4; }
  if (nb) { L->o_nbs = o; o += 32 * SDR_LANES * 16; L->o_mask = o; o += 3 * 128 * SDR_LANES; }
  if (als) { L->o_alsc = o; o += 128 * SDR_LANES * 4; }
  L->ins_row = T + 4; /* staging rows padded by 16 bytes: a lane reading its own row with 16-byte loads is bank-conflict free */
  L->in_depth = T == 32 ? 1 : (T == 16 ? 2 : 4); /* the request has to cover the DRAM latency: about one 32-sample tile time */
  if (in_depth_override > 0 && in_depth_override <= 4) L->in_depth = in_depth_override;
  L->o_ins = o; o += L->in_depth * 2 * SDR_LANES * L->ins_row * 4;
  L->o_outs = o; o += SDR_LANES * L->ins_row * 4;
  const int fixed = o;
  /* rings: grow round robin while the budget allows */
  int *ring[6]; int cost[6]; int n_ring = 0;
  ring[n_ring] = &L->nr; cost[n_ring++] = 2 * tile_b;
  ring[n_ring] = &L->na; cost[n_ring++] = tile_b;
  if (cls == CLS_SSB) { ring[n_ring] = &L->ni; cost[n_ring++] = 2 * tile_b; /* the Q ring grows with it */ }
  else { ring[n_ring] = &L->nz; cost[n_ring++] = 2 * tile_b; ring[n_ring] = &L->nz2; cost[n_ring++] = 2 * tile_b; }
  ring[n_ring] = &L->nc; cost[n_ring++] = tile_b;
  for (int pass = 0; pass < max_slack; pass++) {
    for (int r = 0; r < n_ring; r++) {
      int total = fixed + L->nr * 2 * tile_b + L->na * tile_b + L->nc * tile_b;
      if (cls == CLS_SSB) total += (L->hq_tiles * T / 2 + SDR_HQ_MIRROR) * SDR_LANES * 8 + L->ni * tile_b;
      else total += L->nz * 2 * tile_b + L->nz2 * 2 * tile_b;
      if (total + cost[r] > budget) continue;
      *ring[r] += 1;
      if (cls == CLS_SSB && ring[r] == &L->ni) L->hq_tiles += 1;
    }
  }
  L->o_r = o; o += L->nr * 2 * tile_b;
  if (cls == CLS_SSB) {
    L->hq_rows = L->hq_tiles * T / 2;
    L->o_hq = o; o += (L->hq_rows + SDR_HQ_MIRROR) * SDR_LANES * 8;
    L->o_hi = o; o += L->ni * tile_b;
  } else {
    L->o_z = o; o += L->nz * 2 * tile_b;
    L->o_z2 = o; o += L->nz2 * 2 * tile_b;
  }
  L->o_a = o; o += L->na * tile_b;
  L->o_c = o; o += L->nc * tile_b;
  L->smem_bytes = o;
  }
  /* warps: one per active stage, in stage order until a measured placement is supplied (lay_place).  The 32-sample plan
   * keeps all 14 warps whatever the bucket -- stages it does not need only keep step with the others -- so that the one
   * measured placement of stages on SM sub-partitions serves every bucket. */
  memset(L->prog, 0xFF, sizeof L->prog);
  L->n_warps = 0;
  if (merged) {
    static const uint8_t P[7][4] = {{ST_IN, ST_OUT, 0xFF, 0xFF}, {ST_IFI, 0xFF, 0xFF, 0xFF}, {ST_IFQ, 0xFF, 0xFF, 0xFF}, {ST_PLL, 0xFF, 0xFF, 0xFF},
                                    {ST_NCO2, ST_IMGI, ST_IMGQ, ST_MAG}, {ST_AUD, 0xFF, 0xFF, 0xFF}, {ST_AGC, 0xFF, 0xFF, 0xFF}};
    /* placement: warp id % 4 = SM sub-partition, the higher id has priority there */
    static const uint8_t order_default[7] = {6, 4, 1, 3, 0, 5, 2}; /* measured: tools/map_search.py --cls envmerged */
    const uint8_t *order = merged_order ? merged_order : order_default;
    for (int w = 0; w < 7; w++) memcpy(L->prog[w], P[order[w]], 4);
    L->n_warps = 7;
  } else {
    for (int s = 0; s < SDR_STAGES; s++) if (L->active[s] || T == 32) L->prog[L->n_warps++][0] = (uint8_t)s;
  }
  for (int w = 0; w < L->n_warps; w++) L->stage_of_warp[w] = L->prog[w][0];
  lay_rules(L);
  return lay_check(L);
}

static inline int lay_build(SdrLay *L, int cls, uint32_t feat, int T, int budget, int max_slack) {
  return lay_build_ex(L, cls, feat, T, budget, max_slack, 0, (const uint8_t *)0);
}

/* The ALS + output post-pass (SdrLay::cls == CLS_ALS).  The LMS line enhancer (C:324-352) is the chain's last stage and by far
 * its slowest (one warp sweeps 55 taps for every 4 samples, each sweep waiting for the error of the sweep before), and a
 * group that carries it fills an SM's shared memory, so a handle with more groups than SMs queues in waves whose length the
 * ALS warp sets.  Such a handle's ALS buckets run in two launches instead: the chain up to the AGC -- the bucket's plan
 * without ALS, its output stage writing the AGC output tiles as they are to a scratch plane (SdrLaunch::flags &
 * SDRL_RAW_OUT) -- and this plan: ONE warp per group that fetches the scratch tiles one tile ahead (asynchronous copies
 * straight into the ALS input ring: the scratch plane has the ring's [sample][lane] layout) and runs the chain's ALS +
 * output stage.  A quarter of an SM's shared memory or less per group: four groups share an SM, each on a scheduler of its
 * own.
 *   m_max, reach_max: the largest tap count M and the largest M + delay among the bucket's channels (C:393-398).
 * Tap array: the sweeps load taps up to index M + 4 (five are fetched ahead), so M + 5 rows (at most 128) are kept.
 * Ring: the sweeps for a tile reach back M + delay samples and load (without using) up to 9 more; one more slot holds the
 * tile being swept and one the tile being fetched: never one a sweep touches.  At least 5 slots: the state keeps 128 samples.
 * When it fits the quarter SM, the ring is kept TWICE, back to back (every tile is fetched into slot s and slot s + nc):
 * any run of ring positions is then contiguous in one of the two copies and no sweep ever wraps (als_tile<true>).
 * A single warp in program order: no hand-over rules. */
static inline int lay_build_als(SdrLay *L, int m_max, int reach_max) {
  memset(L, 0, sizeof *L);
  const int T = 32, tile_b = T * SDR_LANES * 4;
  if (m_max < 0) m_max = 0; if (m_max > 128) m_max = 128;
  if (reach_max < m_max) reach_max = m_max; if (reach_max > 129) reach_max = 129;
  L->cls = CLS_ALS; L->feat = LF_ALS; L->T = T; L->tpb = 4; L->tpb_sh = 2; L->tile_f = T * SDR_LANES;
  L->active[ST_IN] = L->active[ST_OUT] = 1;
  for (int s = 0; s < 16; s++) { L->bar_of[s] = (uint8_t)s; L->bar_count[s] = 1; }
  L->nr = L->na = L->ni = L->nz = L->nz2 = L->hq_tiles = 1; L->ins_row = T + 4; L->in_depth = 1;
  L->nc = (reach_max + 9 + T - 1) / T + 2; if (L->nc < 5) L->nc = 5;
  L->als_rows = m_max + 5 > 128 ? 128 : m_max + 5;
  const int fixed = 128 + L->als_rows * SDR_LANES * 4 + SDR_LANES * L->ins_row * 4;
  L->als_mirror = fixed + 2 * L->nc * tile_b <= (233472 - 4 * 1024) / 4;
  int o = 0;
  L->o_cid = o; o += 128;
  L->o_alsc = o; o += L->als_rows * SDR_LANES * 4;
  L->o_outs = o; o += SDR_LANES * L->ins_row * 4;
  L->o_c = o; o += (L->als_mirror ? 2 : 1) * L->nc * tile_b;
  L->smem_bytes = o;
  memset(L->prog, 0xFF, sizeof L->prog);
  L->n_warps = 1; L->prog[0][0] = ST_IN; L->prog[0][1] = ST_OUT; L->stage_of_warp[0] = ST_IN;
  for (int s = 0; s < SDR_STAGES; s++) for (int i = 0; i < SDR_MAX_DEPS; i++) { L->deps[s][i].stage = -1; L->deps[s][i].kind = 0; L->deps[s][i].k = 0; }
  return 0;
}

/* placement: `map` = stage id of physical warp w in bits 4w..4w+3 (as many nibbles as the launch has warps); ignored unless it
 * names every active stage exactly once */
static inline int lay_place(SdrLay *L, unsigned long long map) {
  unsigned seen = 0, want = 0;
  for (int s = 0; s < SDR_STAGES; s++) if (L->active[s] || L->T == 32) want |= 1u << s;
  for (int w = 0; w < L->n_warps; w++) seen |= 1u << ((map >> (4 * w)) & 15);
  if (seen != want) return 1;
  for (int w = 0; w < L->n_warps; w++) if (L->prog[w][1] != 0xFF) return 1; /* merged programs keep their own placement */
  for (int w = 0; w < L->n_warps; w++) { L->stage_of_warp[w] = (uint8_t)((map >> (4 * w)) & 15); L->prog[w][0] = L->stage_of_warp[w]; }
  return 0;
}

#endif
