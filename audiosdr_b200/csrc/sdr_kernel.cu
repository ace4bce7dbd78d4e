/* sdr_kernel.cu -- sm_100a kernels of the batched receiver chain and their launchers.
 *
 * sdr_pipeline_kernel: one CTA per 32-channel group, one warp per pipeline stage of the launch's bucket (see
 * sdr_pipeline.cuh, sdr_lay.h).  Build flags matter for parity: -fmad=false (no FMA contraction; the reference
 * rounds every product and sum separately), IEEE division and square root (nvcc defaults), no FTZ.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "sdr_kernel.h"
#include "sdr_pipeline.cuh"

using namespace sdrk;

__constant__ float2 c_hilbert2[64]; /* compact Hilbert half, H:757-774, every tap twice: the multiplicand pairs of the packed FIR */

/* ---- stage-to-stage hand-over (sdr_lay.h): one mbarrier per (stage, tile mod SDR_BAR_W), arrival count 1.  A stage
 * signals a finished tile with one arrive by lane 0 after re-converging the warp (the warp barrier orders the other
 * lanes' shared-memory and global writes before the arrive, whose release semantics publish them to the waiting
 * stages); a waiting stage polls the barrier's phase parity with mbarrier.try_wait (acquire), which suspends the warp in
 * hardware instead of spinning in the issue slots of the stages that share its scheduler. */
__device__ __forceinline__ void bar_init(uint32_t addr, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}

/* CTA-wide barrier of warps that arrive from different instructions (every warp is a different stage): the unaligned
 * form after re-converging the warp.  Used once per launch, between the stages' state loads and their tile loops. */
__device__ __forceinline__ void cta_barrier() {
  __syncwarp();
  asm volatile("barrier.sync 0;" ::: "memory");
}

template <class Body>
__device__ __forceinline__ void pipeline_loop(const Ctx &x, int stage, Body body) {
  unsigned long long *prof = x.prof ? x.L->prof : nullptr; /* nullptr at compile time in the product kernel */
  const long long t_loaded = prof ? clock64() : 0;
  cta_barrier(); /* histories and tables are in shared memory */
  const uint32_t n = x.L->n_tiles, tpbm = (uint32_t)(x.Y->tpb - 1);
  const uint32_t bars = (uint32_t)__cvta_generic_to_shared(x.smem + x.Y->o_bar);
  const bool skip = prof && ((x.L->diag_skip >> stage) & 1u);
  long long busy = 0, waiting = 0;
  const long long t_begin = prof ? clock64() : 0;
#pragma unroll 1
  for (uint32_t t = 0; t < n; t++) {
    const long long tw = prof ? clock64() : 0;
#pragma unroll 1
    for (int i = 0; i < SDR_MAX_DEPS; i++) {
      const SdrDep d = x.Y->deps[stage][i];
      if (d.stage < 0) break;
      const long long u = d.kind ? (long long)(t | tpbm) : (long long)t + d.k;
      if (u < 0) continue;
      bar_wait(bars + (uint32_t)(d.stage * SDR_BAR_W + ((uint32_t)u & (SDR_BAR_W - 1))) * 8u, ((uint32_t)u / SDR_BAR_W) & 1u);
    }
    const long long t0 = prof ? clock64() : 0;
    if (!skip) body(t);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) bar_arrive(bars + (uint32_t)(stage * SDR_BAR_W + (t & (SDR_BAR_W - 1))) * 8u);
    if (prof) { const long long t1 = clock64(); waiting += t0 - tw; busy += t1 - t0; }
  }
  if (prof && (threadIdx.x & 31) == 0) {
    unsigned long long *row = prof + (size_t)blockIdx.x * SDR_PROF_SLOTS;
    row[stage] += (unsigned long long)busy;
    row[16 + stage] += (unsigned long long)waiting;
    if (x.Y->stage_of_warp[0] == stage) { row[14] += (unsigned long long)(clock64() - t_begin); row[15] += (unsigned long long)(t_loaded - x.t0); }
  }
}

/* One warp = one stage.  Stages common to both pipeline classes are instantiated once (ring offsets are run-time
 * values of the launch's plan) to keep the kernel's instruction footprint small: every warp runs different code, so
 * the hot loops of all stages have to share the instruction caches. */
__device__ __forceinline__ void run_stage(const Ctx &x, int stage, int lane) {
  const bool ssb = x.Y->cls == CLS_SSB;
  /* every 4-section cascade of the chain (IF rails, audio band-pass, AM image rails) runs through this one site */
  const bool is_if = stage == ST_IFI || stage == ST_IFQ, is_aud = stage == ST_AUD, is_img = !ssb && (stage == ST_IMGI || stage == ST_IMGQ);
  if (is_if || is_aud || is_img) {
    RoleBiquad r; r.load(x, lane, is_if ? 0 : (is_aud ? 1 : 2), is_if ? stage - ST_IFI : (is_img ? stage - ST_IMGI : 0));
    pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); });
    r.save(x);
    return;
  }
  switch (stage) {
    case ST_IN: {
      RoleIn r; r.load(x, lane);
      pipeline_loop(x, stage, [&](uint32_t t) { r.step_a(x, lane, t); __syncwarp(); r.step_b(x, lane, t); });
      r.save(x, lane);
    } break;
    case ST_NB: { RoleNb r; r.load(x, lane); pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); }); r.save(x, lane); } break;
    case ST_ENVL: { RoleEnvl r; r.load(x, lane); pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); }); } break;
    case ST_NBO: { RoleNbo r; r.load(x, lane); pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); }); } break;
    case ST_AGC: { RoleAgc r; r.load(x, lane); pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); }); r.save(x); } break;
    case ST_OUT: {
      RoleOut r; r.load(x, lane);
      pipeline_loop(x, stage, [&](uint32_t t) { r.step_a(x, lane, t); __syncwarp(); r.step_b(x, lane, t); __syncwarp(); });
      r.save(x, lane);
    } break;
    default:
      if (ssb) {
        if (stage == ST_NCO) {
          RoleNco r; r.load(x, lane);
          { /* does every active lane run the same oscillator? then one table per tile serves the whole group */
            const unsigned act = __ballot_sync(0xffffffffu, r.cid >= 0);
            const int leader = act ? __ffs(act) - 1 : 0;
            const uint32_t pb = __shfl_sync(0xffffffffu, f2u(r.phase), leader), ib = __shfl_sync(0xffffffffu, f2u(r.inc), leader);
            r.uniform = __all_sync(0xffffffffu, r.cid < 0 || (f2u(r.phase) == pb && f2u(r.inc) == ib)) != 0;
            if (r.uniform) { r.phase = u2f(pb); r.inc = u2f(ib); }
          }
          pipeline_loop(x, stage, [&](uint32_t t) {
            if (r.uniform) { r.table_step(x, lane); __syncwarp(); r.mix_step(x, lane, t); }
            else r.step(x, lane, t);
          });
          r.save(x);
        } else {
          const int sub = stage - ST_HIL0;
          RoleHilbert r; r.load(x, lane, sub);
          pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, reinterpret_cast<const float *>(c_hilbert2), lane, sub, t); });
          r.save(x, lane, sub);
        }
      } else {
        if (stage == ST_PLL) { RolePll r; r.load(x, lane); pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); }); r.save(x); }
        else if (stage == ST_NCO2) { RoleNco2 r; r.load(x, lane); pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); }); r.save(x); }
        else { RoleMag r; r.load(x, lane); pipeline_loop(x, stage, [&](uint32_t t) { r.step(x, lane, t); }); r.save(x); }
      }
      break;
  }
}

template <bool PROF>
__device__ __forceinline__ void pipeline_cta(const SdrLaunch &L, unsigned char *smem) {
  Ctx x;
  x.L = &L;
  x.Y = &L.lay;
  x.G = &L.groups[blockIdx.x];
  x.smem = smem;
  x.gidx = (int)blockIdx.x;
  x.prof = PROF;
  x.t0 = PROF ? clock64() : 0;
  const int nthr = (int)blockDim.x;
  for (int i = threadIdx.x; i < 257; i += nthr) x.f(x.Y->o_sine)[i] = L.tabs->sine[i];
  if (threadIdx.x < SDR_LANES) reinterpret_cast<int *>(smem + x.Y->o_cid)[threadIdx.x] = x.G->cid[threadIdx.x];
  {
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem + x.Y->o_bar);
    for (int i = threadIdx.x; i < SDR_STAGES * SDR_BAR_W; i += nthr) bar_init(bars + 8u * (uint32_t)i, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads(); /* stage IN requests its first tile from load(), which needs the channel ids */
  for (int i = threadIdx.x; i < SDR_LUT_SLOTS * SDR_AGC_LUT_STRIDE; i += nthr) { /* the group's AGC tables */
    const int id = x.G->lut_ids[i / SDR_AGC_LUT_STRIDE];
    if (id >= 0) x.f(x.Y->o_lut)[i] = L.agc_luts[(size_t)id * SDR_AGC_LUT_STRIDE + i % SDR_AGC_LUT_STRIDE];
  }
  /* Physical warp -> stage (SdrLay::stage_of_warp).  The warp scheduler of an SM sub-partition favours the HIGHER warp id
   * among eligible warps, and warp id % 4 picks the sub-partition, so the placement decides which stages compete for one
   * scheduler and who wins; the defaults for the 14-stage launches (sdr_types.h) were found by measurement
   * (tools/map_search.py). */
  const int phys = threadIdx.x >> 5, lane = threadIdx.x & 31;
  run_stage(x, (int)x.Y->stage_of_warp[phys], lane);
}

/* The product kernel and its diagnostics twin (per-stage busy / waiting counters, sub-phase timers, stage skipping): the
 * twin is launched only when the handle was created with SDR_ROLE_PROFILE=1.  Keeping the counters out of the product
 * kernel is not cosmetic: its stages are different instruction streams whose loops must share the instruction caches. */
extern "C" __global__ void __launch_bounds__(SDR_THREADS_MAX, 1) sdr_pipeline_kernel(const __grid_constant__ SdrLaunch L) {
  extern __shared__ __align__(16) unsigned char smem[];
  pipeline_cta<false>(L, smem);
}
extern "C" __global__ void __launch_bounds__(SDR_THREADS_MAX, 1) sdr_pipeline_prof_kernel(const __grid_constant__ SdrLaunch L) {
  extern __shared__ __align__(16) unsigned char smem[];
  pipeline_cta<true>(L, smem);
}

/* Zero (or re-seed) state words of listed channels: the side effects of the reference setters that
 * re-initialise filter state / rings.  One thread per (entry, word). */
extern "C" __global__ void sdr_reset_kernel(float *state, unsigned long long ch_stride, const uint32_t *chan, const uint32_t *mask,
                                            uint32_t n) {
  const uint32_t e = blockIdx.y;
  if (e >= n) return;
  const uint32_t c = chan[e], m = mask[e];
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < SDR_STATE_WORDS; w += gridDim.x * blockDim.x) {
    bool z = false;
    if ((m & SDRK_R_IF) && w >= W_IF_I && w < W_IF_Q + 16) z = true;
    if ((m & SDRK_R_IMG) && w >= W_IMG_I && w < W_IMG_Q + 16) z = true;
    if ((m & SDRK_R_AUD) && w >= W_AUD && w < W_AUD + 16) z = true;
    if ((m & SDRK_R_ALS) && w >= W_ALS_C && w < W_ALS_H + 128) z = true;
    if ((m & SDRK_R_NB) && w >= W_NB_MASK && w < W_NB_RING) z = true;
    if (z) state[(size_t)w * ch_stride + c] = 0.0f;
  }
  if (m & SDRK_R_NB) { /* the ring planes are float4 groups: group f of channel c is float4 number f*ch_stride + c */
    float4 *ring = reinterpret_cast<float4 *>(state + (size_t)W_NB_RING * ch_stride);
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < 288; f += gridDim.x * blockDim.x)
      ring[(size_t)f * ch_stride + c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

/* state word `w` of every channel := v (used once at create for non-zero power-on values) */
extern "C" __global__ void sdr_fill_word_kernel(float *state, unsigned long long ch_stride, uint32_t w, float v, uint32_t n_ch) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_ch) state[(size_t)w * ch_stride + c] = v;
}

/* gather `n_words` listed state words of `n` listed channels into out[n][n_words] */
extern "C" __global__ void sdr_gather_kernel(const float *state, unsigned long long ch_stride, const uint32_t *chan, uint32_t n,
                                             const uint32_t *words, uint32_t n_words, float *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n_words) return;
  uint32_t e = i / n_words, k = i % n_words;
  uint32_t c = chan ? chan[e] : e;
  out[i] = state[(size_t)words[k] * ch_stride + c];
}

extern "C" int sdrk_setup_device(const float *hilbert64) {
  float2 h2[64];
  for (int i = 0; i < 64; i++) h2[i] = make_float2(hilbert64[i], hilbert64[i]);
  cudaError_t e = cudaMemcpyToSymbol(c_hilbert2, h2, sizeof h2);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(sdr_pipeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(sdr_pipeline_prof_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return (int)e;
  /* launches that need less than half of an SM's shared memory are meant to share the SM: keep the carve-out at its maximum */
  e = cudaFuncSetAttribute(sdr_pipeline_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(sdr_pipeline_prof_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  return (int)e;
}

extern "C" int sdrk_launch_pipeline(const SdrLaunch *L, void *stream) {
  if (L->n_groups == 0) return 0;
  const int threads = L->lay.n_warps * 32, smem = L->lay.smem_bytes;
  if (L->prof) sdr_pipeline_prof_kernel<<<L->n_groups, threads, smem, (cudaStream_t)stream>>>(*L);
  else sdr_pipeline_kernel<<<L->n_groups, threads, smem, (cudaStream_t)stream>>>(*L);
  return (int)cudaGetLastError();
}

extern "C" int sdrk_launch_reset(float *state, unsigned long long ch_stride, const uint32_t *chan, const uint32_t *mask, uint32_t n,
                                 void *stream) {
  if (n == 0) return 0;
  dim3 grid(4, n);
  sdr_reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(state, ch_stride, chan, mask, n);
  return (int)cudaGetLastError();
}

extern "C" int sdrk_launch_fill_word(float *state, unsigned long long ch_stride, uint32_t w, float v, uint32_t n_ch, void *stream) {
  sdr_fill_word_kernel<<<(n_ch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(state, ch_stride, w, v, n_ch);
  return (int)cudaGetLastError();
}

extern "C" int sdrk_launch_gather(const float *state, unsigned long long ch_stride, const uint32_t *chan, uint32_t n,
                                  const uint32_t *words, uint32_t n_words, float *out, void *stream) {
  uint32_t tot = n * n_words;
  if (tot == 0) return 0;
  sdr_gather_kernel<<<(tot + 255) / 256, 256, 0, (cudaStream_t)stream>>>(state, ch_stride, chan, n, words, n_words, out);
  return (int)cudaGetLastError();
}

/* ---- FP32 pipe microbenchmark: the denominator of the FP32 roofline (MEASURED_PEAKS.json has no FP32 entry).
 * kind 0: dependent FFMA chains (2 flop / instruction, what a contracting build could reach);
 * kind 1: alternating FMUL / FADD chains (1 flop / instruction: the instruction mix this parity-exact build issues).
 * 8 independent chains per thread hide the 4-cycle pipe latency. */
template <int KIND>
__global__ void __launch_bounds__(256) sdr_fp32_peak_kernel(float *out, int iters, float a, float b) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = a + (float)(threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (KIND == 0) v[i] = __fmaf_rn(v[i], a, b);
        else if (KIND == 1) { v[i] = __fmul_rn(v[i], a); v[i] = __fadd_rn(v[i], b); }
        else v[i] = (float)__fma_rn((double)v[i], (double)a, (double)b); /* KIND 2: F2F + DFMA + F2F per element */
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  if (s == 123.456f) out[0] = s; /* keep the chains alive */
}

/* returns instructions/s (FP32-pipe lane-instructions per second) in *lane_ips and elapsed ms; 0 on success */
extern "C" int sdrk_fp32_peak(int kind, int iters, double *lane_ips, float *ms_out) {
  float *d = nullptr;
  if (cudaMalloc(&d, 4) != cudaSuccess) return 1;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) { /* first pass warms up */
    cudaEventRecord(e0);
    if (kind == 0) sdr_fp32_peak_kernel<0><<<grid, block>>>(d, iters, 0.999f, 0.001f);
    else if (kind == 1) sdr_fp32_peak_kernel<1><<<grid, block>>>(d, iters, 0.999f, 0.001f);
    else sdr_fp32_peak_kernel<2><<<grid, block>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return 2; }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = (double)grid * block * (double)iters * 64.0 * (kind == 1 ? 2.0 : 1.0); /* kind 2 counts DFMAs */
  *lane_ips = instr / (ms * 1e-3);
  *ms_out = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return 0;
}

/* ---- self-test: the batched branch-free envelope (sqrt_hack_batch) against the plain IEEE-divide form (sqrt_hack),
 * bit for bit, over `n` float bit patterns starting at `first` with stride `step` (covers every exponent). */
__global__ void sdr_selftest_envelope_kernel(unsigned first, unsigned step, unsigned long long n, unsigned long long *bad) {
  unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 8ull;
  unsigned long long local = 0;
  for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x * 8ull) {
    float x[8], e[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = fabsf(__uint_as_float(first + (unsigned)((i + k) * step)));
    sqrt_hack_batch<8>(x, e);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float w = sqrt_hack(x[k]);
      const bool same = __float_as_uint(w) == __float_as_uint(e[k]) || (w != w && e[k] != e[k]);
      if (!same) local++;
    }
  }
  if (local) atomicAdd(bad, local);
}

/* ---- self-test: div_inrange (the PLL's division without range check and slow-path branch) against the IEEE divide, bit for
 * bit, over `n` pseudo-random operand pairs of both signs whose magnitudes cover [2^-60, 2^60] exponent by exponent
 * (mantissas from a 64-bit mix of the pair index; every 16th pair sits on a power of two or one ulp beside it). */
__global__ void sdr_selftest_divide_kernel(unsigned long long seed, unsigned long long n, unsigned long long *bad) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long local = 0;
  for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + i * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    unsigned ma = (unsigned)z & 0x7FFFFFu, mb = (unsigned)(z >> 23) & 0x7FFFFFu;
    const unsigned ea = 67u + (unsigned)((z >> 46) % 121u), eb = 67u + (unsigned)((z >> 54) % 121u); /* 2^-60 .. 2^60 */
    if ((i & 15ull) == 0) { ma = (z >> 62) & 1 ? 0u : 0x7FFFFFu; mb = (z >> 63) ? 0u : 1u; }
    if (ea == 187u) ma = 0; /* 2^60 itself is the upper end */
    if (eb == 187u) mb = 0;
    const float a = __uint_as_float(((unsigned)(i & 1) << 31) | (ea << 23) | ma);
    const float b = __uint_as_float(((unsigned)((i >> 1) & 1) << 31) | (eb << 23) | mb);
    if (__float_as_uint(div_inrange(a, b)) != __float_as_uint(__fdiv_rn(a, b))) local++;
  }
  if (local) atomicAdd(bad, local);
}

extern "C" int sdrk_selftest_divide(unsigned long long seed, unsigned long long n, unsigned long long *mismatches) {
  unsigned long long *d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return 1;
  cudaMemset(d, 0, 8);
  sdr_selftest_divide_kernel<<<148 * 8, 256>>>(seed, n, d);
  cudaError_t e = cudaMemcpy(mismatches, d, 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : 2;
}

extern "C" int sdrk_selftest_envelope(unsigned first, unsigned step, unsigned long long n, unsigned long long *mismatches) {
  unsigned long long *d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return 1;
  cudaMemset(d, 0, 8);
  sdr_selftest_envelope_kernel<<<148 * 8, 256>>>(first, step, n, d);
  cudaError_t e = cudaMemcpy(mismatches, d, 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : 2;
}
