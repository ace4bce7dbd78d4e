/* sdr_aux.cu -- the two blocks either side of the receiver chain, batched: kernels and the C ABI
 * of include/sdr_aux.h.  Per-lane arithmetic lives in sdr_aux_core.cuh (also run on the host by tests/emu/aux_emu.cpp).
 *
 *  I/Q generator (AudioIQgenerator.cpp:33-87): a 257-tap Hilbert FIR in its 64-product compact form plus a 128-sample delay,
 *    purely feed-forward, so the kernel is parallel over channels AND time: one CTA = one channel x 4096 outputs, the
 *    input span (+256 history) converted to float once into shared memory, one lane = 16 consecutive outputs with two
 *    sliding 16-register windows.  192 unfused FP32 operations per output sample against 6 bytes of traffic: FP32-issue
 *    bound.  History (the last 256 inputs per channel) is the only state, kept in two alternating buffers.
 *  Pre-processor (AudioSDRpreProcessor.cpp:46-138): with the detector off a channel is a one-sample shift and/or a swap --
 *    a feed-forward copy, 8 bytes moved per sample, HBM bound (pp_static_kernel).  Channels whose detector is running are
 *    compacted into a list and handled block by block by pp_detect_kernel (lane = channel, 128-point FFT per block in
 *    shared memory), because their correction may change from one block to the next.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sdr_aux.h"
#include "aux_tables.inc"
#include "sdr_aux_core.cuh"

__constant__ float c_iq_h[64];
__constant__ float c_fft_tw[128];
__constant__ float c_fft_tw256[256];

/* ================================================================== I/Q generator ==== */
#define IQ_SEG 4096
#define IQ_THREADS 256
#define IQ_XS_WORDS (IQ_SEG + 256 + (IQ_SEG + 256) / 16 + 16)

struct IqLaunch {
  const int16_t *x; int16_t *oi, *oq;
  unsigned long long in_pitch, out_pitch;
  const int16_t *hist_in; int16_t *hist_out; /* [n_channels][256] */
  const float2 *gains;                       /* (gainI, gainQ) per channel */
  uint32_t n_samples, n_channels, segs;
};

__device__ __forceinline__ void iq_unpack8(const int4 v, float *f) {
  const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    f[2 * i] = aux_q15_to_float((int)(int16_t)(w[i] & 0xFFFF));
    f[2 * i + 1] = aux_q15_to_float(w[i] >> 16);
  }
}

__global__ void __launch_bounds__(IQ_THREADS) iq_generate_kernel(const IqLaunch L) {
  __shared__ float xs[IQ_XS_WORDS];
  const uint32_t ch = blockIdx.x / L.segs, seg = blockIdx.x % L.segs;
  const long base = (long)seg * IQ_SEG;
  const int16_t *row = L.x + (size_t)ch * L.in_pitch;
  /* the span [base-256, base+SEG) as floats at padded positions; 8 samples (16 bytes) per thread and pass */
  for (int t = threadIdx.x; t < (IQ_SEG + 256) / 8; t += IQ_THREADS) {
    const long n = base - 256 + 8 * t;
    int4 v = make_int4(0, 0, 0, 0);
    if (n < 0) v = *reinterpret_cast<const int4 *>(L.hist_in + (size_t)ch * 256 + (256 + n));
    else if (n < (long)L.n_samples) v = __ldg(reinterpret_cast<const int4 *>(row + n));
    float f[8];
    iq_unpack8(v, f);
    float *dst = xs + 8 * t + ((8 * t) >> 4);
#pragma unroll
    for (int i = 0; i < 8; i++) dst[i] = f[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long n0 = base + 512 * warp + 16 * lane;
  if (n0 >= (long)L.n_samples) return;
  const int q0 = 256 + 512 * warp + 16 * lane;
  float acc[16];
  iq_lane_fir(xs, q0, c_iq_h, acc);
  const float2 g = L.gains[ch];
  uint32_t wi[8], wq[8];
  if (g.x == 1.0f && g.y == 1.0f) { /* the constructor's balance (IQ.h:72-73): uniform over the CTA, no FP64, no range test */
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      const int i0 = aux_to_pcm_unit(iq_lane_delayed(xs, q0, c)), i1 = aux_to_pcm_unit(iq_lane_delayed(xs, q0, c + 1));
      const int q0v = aux_to_pcm_unit(acc[c]), q1v = aux_to_pcm_unit(acc[c + 1]);
      wi[c >> 1] = ((uint32_t)i0 & 0xFFFFu) | ((uint32_t)i1 << 16);
      wq[c >> 1] = ((uint32_t)q0v & 0xFFFFu) | ((uint32_t)q1v << 16);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      const int i0 = aux_to_pcm(iq_lane_delayed(xs, q0, c), g.x), i1 = aux_to_pcm(iq_lane_delayed(xs, q0, c + 1), g.x);
      const int q0v = aux_to_pcm(acc[c], g.y), q1v = aux_to_pcm(acc[c + 1], g.y);
      wi[c >> 1] = ((uint32_t)i0 & 0xFFFFu) | ((uint32_t)i1 << 16);
      wq[c >> 1] = ((uint32_t)q0v & 0xFFFFu) | ((uint32_t)q1v << 16);
    }
  }
  uint4 *di = reinterpret_cast<uint4 *>(L.oi + (size_t)ch * L.out_pitch + n0);
  uint4 *dq = reinterpret_cast<uint4 *>(L.oq + (size_t)ch * L.out_pitch + n0);
  di[0] = make_uint4(wi[0], wi[1], wi[2], wi[3]); di[1] = make_uint4(wi[4], wi[5], wi[6], wi[7]);
  dq[0] = make_uint4(wq[0], wq[1], wq[2], wq[3]); dq[1] = make_uint4(wq[4], wq[5], wq[6], wq[7]);
}

/* the last 256 inputs of every channel become the next call's history (other buffer: no ordering hazard) */
__global__ void iq_history_kernel(const IqLaunch L) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ch = idx >> 5, j = idx & 31;
  if (ch >= L.n_channels) return;
  const long n = (long)L.n_samples - 256 + 8 * (long)j;
  int4 v;
  if (n < 0) v = *reinterpret_cast<const int4 *>(L.hist_in + (size_t)ch * 256 + (256 + n));
  else v = __ldg(reinterpret_cast<const int4 *>(L.x + (size_t)ch * L.in_pitch + n));
  *reinterpret_cast<int4 *>(L.hist_out + (size_t)ch * 256 + 8 * j) = v;
}

/* ================================================================== pre-processor ==== */
struct PpLaunch {
  const int16_t *I, *Q; int16_t *oi, *oq;
  unsigned long long in_pitch, out_pitch;
  PpState *state;
  uint32_t n_channels, n_blocks;
  uint32_t *auto_list, *auto_count;
};

__global__ void pp_apply_kernel(PpState *st, uint32_t n_channels, const PpSetterCall *calls, uint32_t n_calls) {
  const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= n_channels) return;
  PpState s = st[ch];
  for (uint32_t k = 0; k < n_calls; k++) {
    const PpSetterCall c = calls[k];
    if (c.channel == ch || c.channel == 0xFFFFFFFFu) pp_apply(s, c.setter, c.arg);
  }
  st[ch] = s;
}

__global__ void pp_list_kernel(const PpLaunch L) {
  const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= L.n_channels) return;
  if (L.state[ch].autod) L.auto_list[atomicAdd(L.auto_count, 1u)] = ch;
}

__device__ __forceinline__ void unpack_i16x8(const int4 v, int16_t *o) {
  const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) { o[2 * i] = (int16_t)(w[i] & 0xFFFF); o[2 * i + 1] = (int16_t)(w[i] >> 16); }
}
__device__ __forceinline__ int4 pack_i16x8(const int16_t *o) {
  int w[4];
#pragma unroll
  for (int i = 0; i < 4; i++) w[i] = (int)(((uint32_t)(uint16_t)o[2 * i]) | ((uint32_t)(uint16_t)o[2 * i + 1] << 16));
  return make_int4(w[0], w[1], w[2], w[3]);
}

/* detector off: one thread = one 16-byte chunk of both rails.  8 bytes in + 8 bytes out per 2 samples... i.e. 8 B/sample. */
__global__ void __launch_bounds__(256) pp_static_kernel(const PpLaunch L) {
  const uint32_t chunks = L.n_blocks * 16; /* per channel */
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ch = (uint32_t)(idx / chunks), j = (uint32_t)(idx % chunks);
  if (ch >= L.n_channels) return;
  const PpState s = L.state[ch];
  if (s.autod) return; /* pp_detect_kernel owns this channel for the call */
  const int16_t *ri = L.I + (size_t)ch * L.in_pitch, *rq = L.Q + (size_t)ch * L.in_pitch;
  const size_t n = (size_t)j * 8;
  const int4 a = __ldg(reinterpret_cast<const int4 *>(ri + n)), b = __ldg(reinterpret_cast<const int4 *>(rq + n));
  int4 oa = a, ob = b;
  if (s.corr != 0) {
    int16_t vi[8], vq[8], oi[8], oq[8];
    unpack_i16x8(a, vi); unpack_i16x8(b, vq);
    int16_t pi = (int16_t)s.saved, pq = (int16_t)s.saved;
    if (j) { if (s.corr == 1) pi = ri[n - 1]; else pq = rq[n - 1]; }
    pp_static_chunk(vi, vq, pi, pq, (j & 15) == 0, s.corr, 0, oi, oq);
    oa = pack_i16x8(oi); ob = pack_i16x8(oq);
  }
  int4 *di = reinterpret_cast<int4 *>(L.oi + (size_t)ch * L.out_pitch + n), *dq = reinterpret_cast<int4 *>(L.oq + (size_t)ch * L.out_pitch + n);
  *di = s.swap ? ob : oa;
  *dq = s.swap ? oa : ob;
}

/* detector off: savedSample after the call = the delayed rail's last input sample (PP.cpp:62-65 / 67-70) */
__global__ void pp_finish_kernel(const PpLaunch L) {
  const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= L.n_channels) return;
  const PpState s = L.state[ch];
  if (s.autod || s.corr == 0) return;
  const size_t last = (size_t)L.n_blocks * 128 - 1;
  L.state[ch].saved = (s.corr == 1 ? L.I : L.Q)[(size_t)ch * L.in_pitch + last];
}

/* detector running: one CTA = 32 listed channels (lane = channel) x PP_W warps that share each channel's work, block
 * after block.  Rows are staged through shared memory with whole-row (coalesced) transfers; a lane's row is 65 words from
 * its neighbour's (odd: no bank conflicts); the FFT buffer is [256][32] floats, lane-interleaved.  Warp 0 owns the
 * per-channel state (correction, counters); the butterflies of a stage, the conversions and the rows are split over
 * the warps with a CTA barrier between dependent phases. */
#define PP_W 8
#define PP_ROW_WORDS 65
#define PP_DETECT_SMEM (256 * 32 * 4 + 2 * 32 * PP_ROW_WORDS * 4 + 2 * 32 * 4)
__global__ void __launch_bounds__(32 * PP_W) pp_detect_kernel(const PpLaunch L) {
  extern __shared__ __align__(16) unsigned char sm[];
  float *fft = reinterpret_cast<float *>(sm);
  uint32_t *rowI = reinterpret_cast<uint32_t *>(sm + 256 * 32 * 4), *rowQ = rowI + 32 * PP_ROW_WORDS;
  int *sh_det = reinterpret_cast<int *>(rowQ + 32 * PP_ROW_WORDS), *sh_swap = sh_det + 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t n_auto = *L.auto_count, first = blockIdx.x * 32;
  if (first >= n_auto) return;
  const int ch = first + lane < n_auto ? (int)L.auto_list[first + lane] : -1;
  PpState s;
  s.corr = s.saved = s.fail = s.succ = s.swap = s.autod = s.pad0 = s.pad1 = 0;
  if (ch >= 0) s = L.state[ch]; /* every warp reads it; only warp 0's copy evolves and is written back */
  float *mine = fft + lane;
  for (uint32_t b = 0; b < L.n_blocks; b++) {
    for (int r = w; r < 32; r += PP_W) {
      const int chr = __shfl_sync(0xFFFFFFFFu, ch, r);
      if (chr >= 0) {
        const uint32_t *si = reinterpret_cast<const uint32_t *>(L.I + (size_t)chr * L.in_pitch + (size_t)b * 128);
        const uint32_t *sq = reinterpret_cast<const uint32_t *>(L.Q + (size_t)chr * L.in_pitch + (size_t)b * 128);
        rowI[r * PP_ROW_WORDS + lane] = __ldg(si + lane); rowI[r * PP_ROW_WORDS + 32 + lane] = __ldg(si + 32 + lane);
        rowQ[r * PP_ROW_WORDS + lane] = __ldg(sq + lane); rowQ[r * PP_ROW_WORDS + 32 + lane] = __ldg(sq + 32 + lane);
      }
    }
    __syncthreads();
    int16_t *ri = reinterpret_cast<int16_t *>(rowI + lane * PP_ROW_WORDS), *rq = reinterpret_cast<int16_t *>(rowQ + lane * PP_ROW_WORDS);
    if (w == 0) {
      if (ch >= 0) pp_correct_block(ri, rq, s);
      sh_det[lane] = ch >= 0 && s.autod;
      sh_swap[lane] = s.swap;
    }
    __syncthreads();
    const bool det = sh_det[lane] != 0;
    if (__any_sync(0xFFFFFFFFu, det)) { /* same lanes, same flags in every warp: uniform over the CTA */
      if (det) {
        for (int i = w; i < 128; i += PP_W) {
          mine[(2 * i) * 32] = aux_q15_to_float(ri[i]);
          mine[(2 * i + 1) * 32] = aux_q15_to_float(rq[i]);
        }
      }
      __syncthreads();
      for (int st = 0; st < 7; st++) {
        if (det) pp_fft128_stage(mine, 32, c_fft_tw, st, w, PP_W);
        __syncthreads();
      }
      float pw[128 / PP_W];
      if (det) {
#pragma unroll
        for (int k = 0; k < 128 / PP_W; k++) pw[k] = pp_power(mine, 32, w + k * PP_W);
      }
      __syncthreads();
      if (det) {
#pragma unroll
        for (int k = 0; k < 128 / PP_W; k++) mine[(w + k * PP_W) * 32] = pw[k];
      }
      __syncthreads();
      if (w == 0 && det) pp_decide(mine, 32, s);
    }
    for (int r = w; r < 32; r += PP_W) {
      const int chr = __shfl_sync(0xFFFFFFFFu, ch, r), sw = sh_swap[r];
      if (chr >= 0) {
        const uint32_t *a = (sw ? rowQ : rowI) + r * PP_ROW_WORDS, *c = (sw ? rowI : rowQ) + r * PP_ROW_WORDS;
        uint32_t *di = reinterpret_cast<uint32_t *>(L.oi + (size_t)chr * L.out_pitch + (size_t)b * 128);
        uint32_t *dq = reinterpret_cast<uint32_t *>(L.oq + (size_t)chr * L.out_pitch + (size_t)b * 128);
        di[lane] = a[lane]; di[32 + lane] = a[32 + lane];
        dq[lane] = c[lane]; dq[32 + lane] = c[32 + lane];
      }
    }
    __syncthreads();
  }
  if (w == 0 && ch >= 0) L.state[ch] = s;
}

/* ================================================================== grabber ==== */
/* AudioGrabberComplex256.cpp:50-72 over a whole call: the snapshot after the call is the last pair of blocks that completed
 * (global block index odd), its first half possibly carried over from the previous call.  96 threads per channel: 64 write
 * the snapshot (4 complex samples each), 32 the carried half. */
struct GrabLaunch {
  const int16_t *I, *Q; unsigned long long pitch;
  const int16_t *half_in; int16_t *half_out; /* [n_channels][256]: first block of an incomplete pair, interleaved */
  int16_t *outb;                             /* [n_channels][512] */
  uint32_t n_channels, n_blocks, t0_odd;
};

__device__ __forceinline__ int4 grab_interleave4(const int16_t *I, const int16_t *Q) { /* 4 samples of each rail -> re,im,re,im,... */
  const uint2 a = *reinterpret_cast<const uint2 *>(I), b = *reinterpret_cast<const uint2 *>(Q);
  int4 o;
  o.x = (int)__byte_perm(a.x, b.x, 0x5410); o.y = (int)__byte_perm(a.x, b.x, 0x7632);
  o.z = (int)__byte_perm(a.y, b.y, 0x5410); o.w = (int)__byte_perm(a.y, b.y, 0x7632);
  return o;
}

__global__ void grab_update_kernel(const GrabLaunch L) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ch = idx / 96, j = idx % 96;
  if (ch >= L.n_channels) return;
  const int n = (int)L.n_blocks;
  const int16_t *ri = L.I + (size_t)ch * L.pitch, *rq = L.Q + (size_t)ch * L.pitch;
  if (j < 64) {
    /* last block of the call that completes a pair */
    int bs = ((int)L.t0_odd + n - 1) & 1 ? n - 1 : n - 2;
    if (bs < 0) return; /* no pair completed in this call: the snapshot stays */
    const int second = j >= 32, s4 = (int)(j & 31) * 4; /* which block of the pair, first sample of the chunk */
    int4 v;
    if (second) v = grab_interleave4(ri + (size_t)bs * 128 + s4, rq + (size_t)bs * 128 + s4);
    else if (bs >= 1) v = grab_interleave4(ri + (size_t)(bs - 1) * 128 + s4, rq + (size_t)(bs - 1) * 128 + s4);
    else v = *reinterpret_cast<const int4 *>(L.half_in + (size_t)ch * 256 + 2 * s4);
    *reinterpret_cast<int4 *>(L.outb + (size_t)ch * 512 + 8 * j) = v;
  } else {
    const int s4 = (int)(j - 64) * 4;
    int4 v;
    if (((int)L.t0_odd + n) & 1) v = grab_interleave4(ri + (size_t)(n - 1) * 128 + s4, rq + (size_t)(n - 1) * 128 + s4);
    else v = *reinterpret_cast<const int4 *>(L.half_in + (size_t)ch * 256 + 2 * s4); /* not observable; keeps the buffers in step */
    *reinterpret_cast<int4 *>(L.half_out + (size_t)ch * 256 + 2 * s4) = v;
  }
}

/* ================================================================== host side ==== */
static thread_local std::string g_err;
static int fail(int code, const std::string &m) { g_err = m; return code; }
#define CK(call)                                                                                                    \
  do {                                                                                                              \
    cudaError_t e_ = (call);                                                                                        \
    if (e_ != cudaSuccess) return fail(SDR_AUX_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

static int upload_tables() {
  float h[64], tw[128];
  memcpy(h, AUX_IQ_HILBERT, sizeof h);
  memcpy(tw, AUX_FFT_TW, sizeof tw);
  CK(cudaMemcpyToSymbol(c_iq_h, h, sizeof h));
  CK(cudaMemcpyToSymbol(c_fft_tw, tw, sizeof tw));
  float tw256[256];
  memcpy(tw256, AUX_FFT_TW256, sizeof tw256);
  CK(cudaMemcpyToSymbol(c_fft_tw256, tw256, sizeof tw256));
  return 0;
}

static int check_planes(const void *a, const void *b, const void *c, const void *d, size_t in_pitch, size_t out_pitch, uint32_t n_blocks) {
  if (!a || !c || !d || n_blocks == 0) return fail(SDR_AUX_EINVAL, "null plane or n_blocks == 0");
  if ((in_pitch & 7) || (out_pitch & 7) || in_pitch < (size_t)n_blocks * 128 || out_pitch < (size_t)n_blocks * 128)
    return fail(SDR_AUX_EINVAL, "pitch must be a multiple of 8 elements and >= 128*n_blocks");
  if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) return fail(SDR_AUX_EINVAL, "planes must be 16-byte aligned");
  if (c == a || c == b || d == a || d == b || c == d) return fail(SDR_AUX_EINVAL, "output planes must not alias the inputs or each other");
  return 0;
}

/* device staging for the *_process_host entry points */
struct Staging {
  int16_t *p[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t pitch = 0; uint32_t blocks = 0;
  int ensure(uint32_t n_channels, uint32_t n_blocks, int planes) {
    if (n_blocks <= blocks && p[planes - 1]) return 0;
    release();
    pitch = (size_t)n_blocks * 128;
    for (int i = 0; i < planes; i++)
      if (cudaMalloc(&p[i], pitch * n_channels * sizeof(int16_t)) != cudaSuccess) { release(); return fail(SDR_AUX_ENOMEM, "cudaMalloc(staging) failed"); }
    blocks = n_blocks;
    return 0;
  }
  void release() { for (auto &q : p) { if (q) cudaFree(q); q = nullptr; } blocks = 0; }
};

/* ------------------------------------------------------------------ pre-processor ---- */
struct sdr_preproc {
  uint32_t n = 0; int device = 0;
  PpState *state = nullptr;
  uint32_t *list = nullptr, *count = nullptr;
  PpSetterCall *d_calls = nullptr; size_t d_calls_cap = 0;
  std::vector<PpSetterCall> pending;
  bool auto_possible = false; /* some channel may have its detector on */
  uint64_t launches = 0;
  Staging stg;
};

extern "C" int sdr_preproc_create(sdr_preproc_t **out, uint32_t n_channels, int device) {
  if (!out || n_channels == 0) return fail(SDR_AUX_EINVAL, "sdr_preproc_create: bad arguments");
  CK(cudaSetDevice(device));
  if (int rc = upload_tables()) return rc;
  CK(cudaFuncSetAttribute(pp_detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_DETECT_SMEM));
  sdr_preproc *h = new sdr_preproc();
  h->n = n_channels; h->device = device;
  if (cudaMalloc(&h->state, sizeof(PpState) * n_channels) != cudaSuccess || cudaMalloc(&h->list, 4 * (size_t)n_channels) != cudaSuccess ||
      cudaMalloc(&h->count, 4) != cudaSuccess) { sdr_preproc_destroy(h); return fail(SDR_AUX_ENOMEM, "cudaMalloc failed"); }
  CK(cudaMemset(h->state, 0, sizeof(PpState) * n_channels)); /* PP.h:63-72 initialisers: everything 0 / false */
  *out = h;
  return 0;
}

extern "C" void sdr_preproc_destroy(sdr_preproc_t *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->state); cudaFree(h->list); cudaFree(h->count); cudaFree(h->d_calls);
  h->stg.release();
  delete h;
}

extern "C" int sdr_preproc_set(sdr_preproc_t *h, const uint32_t *channels, uint32_t n, uint32_t setter, int32_t arg) {
  if (!h) return fail(SDR_AUX_EINVAL, "null handle");
  if (setter < 1 || setter > 4) return fail(SDR_AUX_EINVAL, "unknown pre-processor setter " + std::to_string(setter));
  if (setter == SDR_PP_setI2SerrorCompensation && (arg < -1 || arg > 1)) return fail(SDR_AUX_EINVAL, "I2S compensation must be -1, 0 or +1");
  if (setter == SDR_PP_startAutoI2SerrorDetection) h->auto_possible = true;
  if (!channels) { h->pending.push_back({0xFFFFFFFFu, setter, arg, 0}); return 0; }
  for (uint32_t i = 0; i < n; i++) if (channels[i] >= h->n) return fail(SDR_AUX_EINVAL, "channel out of range");
  for (uint32_t i = 0; i < n; i++) h->pending.push_back({channels[i], setter, arg, 0});
  return 0;
}

static int pp_flush(sdr_preproc *h, cudaStream_t st) {
  if (h->pending.empty()) return 0;
  if (h->pending.size() > h->d_calls_cap) {
    cudaFree(h->d_calls); h->d_calls = nullptr;
    h->d_calls_cap = h->pending.size() * 2 + 64;
    if (cudaMalloc(&h->d_calls, sizeof(PpSetterCall) * h->d_calls_cap) != cudaSuccess) { h->d_calls_cap = 0; return fail(SDR_AUX_ENOMEM, "cudaMalloc(calls) failed"); }
  }
  /* pageable source: the copy is staged before the call returns, so the vector may be cleared right away */
  CK(cudaMemcpyAsync(h->d_calls, h->pending.data(), sizeof(PpSetterCall) * h->pending.size(), cudaMemcpyHostToDevice, st));
  pp_apply_kernel<<<(h->n + 255) / 256, 256, 0, st>>>(h->state, h->n, h->d_calls, (uint32_t)h->pending.size());
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st)); /* d_calls may be reallocated by the next flush */
  h->pending.clear();
  h->launches++;
  return 0;
}

extern "C" int sdr_preproc_get_status(sdr_preproc_t *h, const uint32_t *channels, uint32_t n, sdr_preproc_status *out) {
  if (!h || !out) return fail(SDR_AUX_EINVAL, "null argument");
  CK(cudaSetDevice(h->device));
  if (int rc = pp_flush(h, 0)) return rc;
  CK(cudaDeviceSynchronize());
  const uint32_t cnt = channels ? n : h->n;
  for (uint32_t i = 0; i < cnt; i++) {
    const uint32_t c = channels ? channels[i] : i;
    if (c >= h->n) return fail(SDR_AUX_EINVAL, "channel out of range");
    PpState s;
    CK(cudaMemcpy(&s, h->state + c, sizeof s, cudaMemcpyDeviceToHost));
    out[i].auto_detect = s.autod; out[i].correction = s.corr; out[i].failure_count = s.fail;
    out[i].success_count = s.succ; out[i].saved_sample = s.saved; out[i].swap = s.swap;
  }
  return 0;
}

extern "C" int sdr_preproc_process_device(sdr_preproc_t *h, const int16_t *I, const int16_t *Q, size_t in_pitch, int16_t *I_out,
                                          int16_t *Q_out, size_t out_pitch, uint32_t n_blocks, void *cuda_stream) {
  if (!h) return fail(SDR_AUX_EINVAL, "null handle");
  if (!Q) return fail(SDR_AUX_EINVAL, "null plane");
  if (int rc = check_planes(I, Q, I_out, Q_out, in_pitch, out_pitch, n_blocks)) return rc;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (int rc = pp_flush(h, st)) return rc;
  PpLaunch L = {I, Q, I_out, Q_out, in_pitch, out_pitch, h->state, h->n, n_blocks, h->list, h->count};
  const uint32_t tb = (h->n + 255) / 256;
  if (h->auto_possible) {
    CK(cudaMemsetAsync(h->count, 0, 4, st));
    pp_list_kernel<<<tb, 256, 0, st>>>(L);
    h->launches++;
  }
  const unsigned long long threads = (unsigned long long)h->n * n_blocks * 16;
  pp_static_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(L);
  pp_finish_kernel<<<tb, 256, 0, st>>>(L);
  h->launches += 2;
  if (h->auto_possible) {
    pp_detect_kernel<<<(h->n + 31) / 32, 32 * PP_W, PP_DETECT_SMEM, st>>>(L);
    h->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

extern "C" int sdr_preproc_process_host(sdr_preproc_t *h, const int16_t *I, const int16_t *Q, size_t in_pitch, int16_t *I_out,
                                        int16_t *Q_out, size_t out_pitch, uint32_t n_blocks) {
  if (!h || !I || !Q || !I_out || !Q_out || n_blocks == 0) return fail(SDR_AUX_EINVAL, "bad arguments");
  CK(cudaSetDevice(h->device));
  if (int rc = h->stg.ensure(h->n, n_blocks, 4)) return rc;
  const size_t w = (size_t)n_blocks * 128 * 2, dp = h->stg.pitch * 2;
  CK(cudaMemcpy2DAsync(h->stg.p[0], dp, I, in_pitch * 2, w, h->n, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpy2DAsync(h->stg.p[1], dp, Q, in_pitch * 2, w, h->n, cudaMemcpyHostToDevice, 0));
  if (int rc = sdr_preproc_process_device(h, h->stg.p[0], h->stg.p[1], h->stg.pitch, h->stg.p[2], h->stg.p[3], h->stg.pitch, n_blocks, nullptr)) return rc;
  CK(cudaMemcpy2DAsync(I_out, out_pitch * 2, h->stg.p[2], dp, w, h->n, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpy2DAsync(Q_out, out_pitch * 2, h->stg.p[3], dp, w, h->n, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return 0;
}

extern "C" uint64_t sdr_preproc_launch_count(const sdr_preproc_t *h) { return h ? h->launches : 0; }

/* ------------------------------------------------------------------ I/Q generator ---- */
struct sdr_iqgen {
  uint32_t n = 0; int device = 0;
  int16_t *hist[2] = {nullptr, nullptr};
  int cur = 0;
  float2 *gains = nullptr;
  std::vector<float2> h_gains;
  bool gains_dirty = true;
  uint64_t launches = 0;
  Staging stg;
};

extern "C" int sdr_iqgen_create(sdr_iqgen_t **out, uint32_t n_channels, int device) {
  if (!out || n_channels == 0) return fail(SDR_AUX_EINVAL, "sdr_iqgen_create: bad arguments");
  CK(cudaSetDevice(device));
  if (int rc = upload_tables()) return rc;
  sdr_iqgen *h = new sdr_iqgen();
  h->n = n_channels; h->device = device;
  h->h_gains.assign(n_channels, make_float2(1.0f, 1.0f)); /* IQ.h:72-73 */
  if (cudaMalloc(&h->hist[0], 512 * (size_t)n_channels) != cudaSuccess || cudaMalloc(&h->hist[1], 512 * (size_t)n_channels) != cudaSuccess ||
      cudaMalloc(&h->gains, sizeof(float2) * n_channels) != cudaSuccess) { sdr_iqgen_destroy(h); return fail(SDR_AUX_ENOMEM, "cudaMalloc failed"); }
  CK(cudaMemset(h->hist[0], 0, 512 * (size_t)n_channels)); /* the static buffers start at zero, IQ.cpp:37-38 */
  CK(cudaMemset(h->hist[1], 0, 512 * (size_t)n_channels));
  *out = h;
  return 0;
}

extern "C" void sdr_iqgen_destroy(sdr_iqgen_t *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->hist[0]); cudaFree(h->hist[1]); cudaFree(h->gains);
  h->stg.release();
  delete h;
}

extern "C" int sdr_iqgen_set_gain_balance(sdr_iqgen_t *h, const uint32_t *channels, uint32_t n, float balance) {
  if (!h) return fail(SDR_AUX_EINVAL, "null handle");
  const float2 g = make_float2(balance, (float)(1.0 / (double)balance)); /* IQ.h:58-59 */
  if (!channels) { for (auto &x : h->h_gains) x = g; }
  else {
    for (uint32_t i = 0; i < n; i++) if (channels[i] >= h->n) return fail(SDR_AUX_EINVAL, "channel out of range");
    for (uint32_t i = 0; i < n; i++) h->h_gains[channels[i]] = g;
  }
  h->gains_dirty = true;
  return 0;
}

extern "C" int sdr_iqgen_process_device(sdr_iqgen_t *h, const int16_t *X, size_t in_pitch, int16_t *I_out, int16_t *Q_out,
                                        size_t out_pitch, uint32_t n_blocks, void *cuda_stream) {
  if (!h) return fail(SDR_AUX_EINVAL, "null handle");
  if (int rc = check_planes(X, nullptr, I_out, Q_out, in_pitch, out_pitch, n_blocks)) return rc;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (h->gains_dirty) {
    CK(cudaMemcpyAsync(h->gains, h->h_gains.data(), sizeof(float2) * h->n, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st)); /* pageable source may change after we return */
    h->gains_dirty = false;
  }
  IqLaunch L;
  L.x = X; L.oi = I_out; L.oq = Q_out; L.in_pitch = in_pitch; L.out_pitch = out_pitch;
  L.hist_in = h->hist[h->cur]; L.hist_out = h->hist[h->cur ^ 1]; L.gains = h->gains;
  L.n_samples = n_blocks * 128; L.n_channels = h->n; L.segs = (L.n_samples + IQ_SEG - 1) / IQ_SEG;
  const unsigned long long grid = (unsigned long long)h->n * L.segs;
  if (grid > 0x7FFFFFFFull) return fail(SDR_AUX_EINVAL, "too many channel-segments for one launch");
  iq_generate_kernel<<<(unsigned)grid, IQ_THREADS, 0, st>>>(L);
  iq_history_kernel<<<(h->n * 32 + 255) / 256, 256, 0, st>>>(L);
  CK(cudaGetLastError());
  h->cur ^= 1;
  h->launches += 2;
  return 0;
}

extern "C" int sdr_iqgen_process_host(sdr_iqgen_t *h, const int16_t *X, size_t in_pitch, int16_t *I_out, int16_t *Q_out,
                                      size_t out_pitch, uint32_t n_blocks) {
  if (!h || !X || !I_out || !Q_out || n_blocks == 0) return fail(SDR_AUX_EINVAL, "bad arguments");
  CK(cudaSetDevice(h->device));
  if (int rc = h->stg.ensure(h->n, n_blocks, 3)) return rc;
  const size_t w = (size_t)n_blocks * 128 * 2, dp = h->stg.pitch * 2;
  CK(cudaMemcpy2DAsync(h->stg.p[0], dp, X, in_pitch * 2, w, h->n, cudaMemcpyHostToDevice, 0));
  if (int rc = sdr_iqgen_process_device(h, h->stg.p[0], h->stg.pitch, h->stg.p[1], h->stg.p[2], h->stg.pitch, n_blocks, nullptr)) return rc;
  CK(cudaMemcpy2DAsync(I_out, out_pitch * 2, h->stg.p[1], dp, w, h->n, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpy2DAsync(Q_out, out_pitch * 2, h->stg.p[2], dp, w, h->n, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return 0;
}

extern "C" uint64_t sdr_iqgen_launch_count(const sdr_iqgen_t *h) { return h ? h->launches : 0; }

/* Spectrum tap on the snapshots (the panadapter that consumes grab()): per channel the 256-point forward
 * complex FFT of the snapshot's (re, im) int16 pairs taken as floats, then the power re^2 + im^2 of every bin in natural bin
 * order.  One CTA = one channel, 128 threads = the 128 butterflies of a stage; the operation network is the radix-2
 * decimation-in-frequency one the test suite's checker restates (aux_cfft256_forward), one rounding per operation. */
__global__ void __launch_bounds__(128) grab_spectrum_kernel(const int16_t *snap, const uint32_t *channels, float *power) {
  __shared__ float2 buf[256];
  const uint32_t ch = channels ? channels[blockIdx.x] : blockIdx.x;
  const int t = threadIdx.x;
  const int2 raw = reinterpret_cast<const int2 *>(snap + (size_t)ch * 512)[t]; /* samples 2t, 2t+1: (re, im, re, im) */
  buf[2 * t] = make_float2((float)(int16_t)(raw.x & 0xFFFF), (float)(int16_t)(raw.x >> 16));
  buf[2 * t + 1] = make_float2((float)(int16_t)(raw.y & 0xFFFF), (float)(int16_t)(raw.y >> 16));
  __syncthreads();
#pragma unroll 1
  for (int half = 128; half >= 1; half >>= 1) {
    const int step = 128 / half, j = t & (half - 1), ia = ((t / half) * 2 * half) + j, ib = ia + half;
    const float2 a = buf[ia], b = buf[ib];
    const float tr = a.x - b.x, ti = a.y - b.y;
    const float wr = c_fft_tw256[2 * j * step], wi = c_fft_tw256[2 * j * step + 1];
    const float p0 = tr * wr, p1 = ti * wi, p2 = tr * wi, p3 = ti * wr;
    buf[ia] = make_float2(a.x + b.x, a.y + b.y);
    buf[ib] = make_float2(p0 - p1, p2 + p3);
    __syncthreads();
  }
#pragma unroll
  for (int k = t; k < 256; k += 128) {
    const float2 v = buf[__brev((unsigned)k) >> 24]; /* 8-bit reversal: natural bin order */
    const float a = v.x * v.x, b = v.y * v.y;
    power[(size_t)blockIdx.x * 256 + k] = a + b;
  }
}

/* ------------------------------------------------------------------ grabber ---- */
struct sdr_grabber {
  uint32_t n = 0; int device = 0;
  int16_t *half[2] = {nullptr, nullptr}, *outb = nullptr;
  int cur = 0;
  uint64_t blocks = 0;     /* update() calls so far (every channel advances together) */
  bool valid = false;      /* _dataBufferValid */
  std::vector<uint8_t> fresh; /* _newDataIsAvailable per channel */
};

extern "C" int sdr_grabber_create(sdr_grabber_t **out, uint32_t n_channels, int device) {
  if (!out || n_channels == 0) return fail(SDR_AUX_EINVAL, "sdr_grabber_create: bad arguments");
  CK(cudaSetDevice(device));
  if (int rc = upload_tables()) return rc; /* the spectrum tap's twiddles */
  sdr_grabber *h = new sdr_grabber();
  h->n = n_channels; h->device = device; h->fresh.assign(n_channels, 0);
  if (cudaMalloc(&h->half[0], 512 * (size_t)n_channels) != cudaSuccess || cudaMalloc(&h->half[1], 512 * (size_t)n_channels) != cudaSuccess ||
      cudaMalloc(&h->outb, 1024 * (size_t)n_channels) != cudaSuccess) { sdr_grabber_destroy(h); return fail(SDR_AUX_ENOMEM, "cudaMalloc failed"); }
  CK(cudaMemset(h->half[0], 0, 512 * (size_t)n_channels));
  CK(cudaMemset(h->half[1], 0, 512 * (size_t)n_channels));
  CK(cudaMemset(h->outb, 0, 1024 * (size_t)n_channels));
  *out = h;
  return 0;
}

extern "C" void sdr_grabber_destroy(sdr_grabber_t *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->half[0]); cudaFree(h->half[1]); cudaFree(h->outb);
  delete h;
}

extern "C" int sdr_grabber_process_device(sdr_grabber_t *h, const int16_t *I, const int16_t *Q, size_t pitch, uint32_t n_blocks, void *cuda_stream) {
  if (!h || !I || !Q || n_blocks == 0) return fail(SDR_AUX_EINVAL, "bad arguments");
  if ((pitch & 7) || pitch < (size_t)n_blocks * 128 || (((uintptr_t)I | (uintptr_t)Q) & 15)) return fail(SDR_AUX_EINVAL, "planes must be 16-byte aligned, pitch a multiple of 8 and >= 128*n_blocks");
  CK(cudaSetDevice(h->device));
  GrabLaunch L = {I, Q, pitch, h->half[h->cur], h->half[h->cur ^ 1], h->outb, h->n, n_blocks, (uint32_t)(h->blocks & 1)};
  const unsigned long long threads = (unsigned long long)h->n * 96;
  grab_update_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(L);
  CK(cudaGetLastError());
  const bool completed = ((h->blocks & 1) + n_blocks) >= 2; /* some block of this call had an odd global index */
  h->cur ^= 1;
  h->blocks += n_blocks;
  if (completed) { h->valid = true; std::fill(h->fresh.begin(), h->fresh.end(), (uint8_t)1); }
  return 0;
}

extern "C" int sdr_grabber_new_data_available(sdr_grabber_t *h, uint32_t channel) {
  if (!h || channel >= h->n) return fail(SDR_AUX_EINVAL, "bad arguments");
  return h->fresh[channel];
}

extern "C" int sdr_grabber_grab(sdr_grabber_t *h, const uint32_t *channels, uint32_t n, int16_t *dest) {
  if (!h || !dest) return fail(SDR_AUX_EINVAL, "bad arguments");
  const uint32_t cnt = channels ? n : h->n;
  for (uint32_t i = 0; i < cnt; i++) if ((channels ? channels[i] : i) >= h->n) return fail(SDR_AUX_EINVAL, "channel out of range");
  CK(cudaSetDevice(h->device));
  int written = 0;
  if (h->valid) { /* AudioGrabberComplex256.cpp:83-88 */
    CK(cudaDeviceSynchronize());
    if (!channels) { CK(cudaMemcpy(dest, h->outb, 1024 * (size_t)h->n, cudaMemcpyDeviceToHost)); }
    else for (uint32_t i = 0; i < cnt; i++) CK(cudaMemcpy(dest + (size_t)i * 512, h->outb + (size_t)channels[i] * 512, 1024, cudaMemcpyDeviceToHost));
    written = (int)cnt;
  }
  for (uint32_t i = 0; i < cnt; i++) h->fresh[channels ? channels[i] : i] = 0; /* AudioGrabberComplex256.cpp:90 */
  return written;
}

extern "C" int sdr_grabber_grab_device(sdr_grabber_t *h, int16_t *dest, void *cuda_stream) {
  if (!h || !dest) return fail(SDR_AUX_EINVAL, "bad arguments");
  CK(cudaSetDevice(h->device));
  if (h->valid) CK(cudaMemcpyAsync(dest, h->outb, 1024 * (size_t)h->n, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
  std::fill(h->fresh.begin(), h->fresh.end(), (uint8_t)0);
  return h->valid ? 1 : 0;
}

/* power[i*256 + k] = |FFT256(snapshot of channel i)|^2[k], device memory, for the listed channels (device array, NULL: all);
 * returns 0 when no snapshot is valid yet (nothing written), 1 when written.  Does not touch the new-data flags. */
extern "C" int sdr_grabber_spectrum_device(sdr_grabber_t *h, const uint32_t *d_channels, uint32_t n, float *d_power, void *cuda_stream) {
  if (!h || !d_power) return fail(SDR_AUX_EINVAL, "bad arguments");
  CK(cudaSetDevice(h->device));
  if (!h->valid) return 0;
  const uint32_t cnt = d_channels ? n : h->n;
  if (cnt == 0) return 1;
  grab_spectrum_kernel<<<cnt, 128, 0, (cudaStream_t)cuda_stream>>>(h->outb, d_channels, d_power);
  CK(cudaGetLastError());
  return 1;
}

/* the same to HOST memory for the listed channels (NULL: all) */
extern "C" int sdr_grabber_spectrum(sdr_grabber_t *h, const uint32_t *channels, uint32_t n, float *power) {
  if (!h || !power) return fail(SDR_AUX_EINVAL, "bad arguments");
  const uint32_t cnt = channels ? n : h->n;
  for (uint32_t i = 0; i < cnt; i++) if ((channels ? channels[i] : i) >= h->n) return fail(SDR_AUX_EINVAL, "channel out of range");
  CK(cudaSetDevice(h->device));
  if (!h->valid) return 0;
  if (cnt == 0) return 0;
  uint32_t *d_ch = nullptr; float *d_p = nullptr;
  int rc = 0;
  do {
    if (cudaMalloc(&d_p, (size_t)cnt * 1024) != cudaSuccess) { rc = fail(SDR_AUX_ENOMEM, "cudaMalloc failed"); break; }
    if (channels) {
      if (cudaMalloc(&d_ch, (size_t)cnt * 4) != cudaSuccess) { rc = fail(SDR_AUX_ENOMEM, "cudaMalloc failed"); break; }
      if (cudaMemcpy(d_ch, channels, (size_t)cnt * 4, cudaMemcpyHostToDevice) != cudaSuccess) { rc = fail(SDR_AUX_ECUDA, "copy failed"); break; }
    }
    grab_spectrum_kernel<<<cnt, 128>>>(h->outb, d_ch, d_p);
    if (cudaGetLastError() != cudaSuccess || cudaMemcpy(power, d_p, (size_t)cnt * 1024, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = fail(SDR_AUX_ECUDA, "spectrum kernel failed"); break; }
    rc = (int)cnt;
  } while (0);
  cudaFree(d_ch); cudaFree(d_p);
  return rc;
}

extern "C" const char *sdr_aux_last_error(void) { return g_err.c_str(); }
extern "C" const char *sdr_aux_version(void) { return "audiosdr_b200 aux 0.1 (sm_100a)"; }
