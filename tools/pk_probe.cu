/* tools/pk_probe.cu -- diagnostics: cost of one Hilbert pair-tap  acc += h * (A - B)  (two outputs) in four instruction forms,
 * registers only, at 1, 2 and 4 warps per SM sub-partition.  Prints cycles per pair-tap per sub-partition.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/pk_probe tools/pk_probe.cu
 *  form 0: 6 scalar instructions (FADD, FADD, FMUL, FMUL, FADD, FADD)
 *  form 1: 3 FFMA2 with run-time constants: fma(B,-1,A), fma(h,d,-0), fma(p,1,acc)
 *  form 2: FADD2 (sub), FMUL2, two scalar FADD
 *  form 3: FADD2 (sub), FMUL2, FFMA2(p, 1 (run time), acc)
 *  form 4: FADD2 (sub), FFMA2(h, d, -0 (run time)), FADD2 (acc + p) */
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int FORM>
__global__ void __launch_bounds__(512) probe(float *out, int iters, const float *k) {
  const float a = k[6], b = k[7];
  const u64 one = pk(k[0], k[1]), mone = pk(k[2], k[3]), mzero = pk(k[4], k[5]);
  u64 acc[4], RA[8], RB[8], H[8];
  float sacc[8];
#pragma unroll
  for (int i = 0; i < 4; i++) { acc[i] = pk(a + i + threadIdx.x, b + i); sacc[2 * i] = a + i; sacc[2 * i + 1] = b - i; }
#pragma unroll
  for (int i = 0; i < 8; i++) { RA[i] = pk(a * (i + 1), b + i); RB[i] = pk(b * (i + 2), a - i); H[i] = pk(0.001f * (i + 1), 0.001f * (i + 1)); }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const u64 A = RA[(r - kk) & 7], B = RB[(r + kk + 1) & 7], h = H[kk];
        if (FORM == 0) {
          float al, ah, bl, bh, hl, hh; unpk(A, al, ah); unpk(B, bl, bh); unpk(h, hl, hh);
          sacc[2 * r] = __fadd_rn(sacc[2 * r], __fmul_rn(hl, __fadd_rn(al, -bl)));
          sacc[2 * r + 1] = __fadd_rn(sacc[2 * r + 1], __fmul_rn(hh, __fadd_rn(ah, -bh)));
        } else if (FORM == 1) {
          acc[r] = fma2(fma2(h, fma2(B, mone, A), mzero), one, acc[r]);
        } else if (FORM == 2) {
          float pl, ph; unpk(mul2(h, sub2(A, B)), pl, ph);
          sacc[2 * r] = __fadd_rn(sacc[2 * r], pl); sacc[2 * r + 1] = __fadd_rn(sacc[2 * r + 1], ph);
        } else if (FORM == 3) {
          acc[r] = fma2(mul2(h, sub2(A, B)), one, acc[r]);
        } else {
          acc[r] = add2(acc[r], fma2(h, sub2(A, B), mzero));
        }
      }
    }
    /* every window element changes every iteration (by run-time constants): nothing is hoistable.  16 extra packed FMAs
     * per 32 pair-taps in every form. */
#pragma unroll
    for (int i = 0; i < 8; i++) { RA[i] = fma2(RA[i], one, mzero); RB[i] = fma2(RB[i], one, mzero); }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) { float lo, hi; unpk(acc[i], lo, hi); s += lo + hi + sacc[2 * i] + sacc[2 * i + 1]; }
  if (s == 123.456f) out[0] = s;
}

template <int FORM>
static void run(int threads, const float *dk, float *dout, int sms) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); probe<FORM><<<sms, threads>>>(dout, iters, dk); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * khz * 1e3;
  const double warps_per_smsp = threads / 128.0;
  printf("form %d, %.0f warp(s) per sub-partition: %.2f cycles per pair-tap per warp, %.2f per sub-partition\n", FORM, warps_per_smsp,
         cycles / (iters * 32.0), cycles / (iters * 32.0 * warps_per_smsp));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const float hk[8] = {1.f, 1.f, -1.f, -1.f, -0.f, -0.f, 0.999f, 0.001f};
  float *dk, *dout; cudaMalloc(&dk, sizeof hk); cudaMalloc(&dout, 4); cudaMemcpy(dk, hk, sizeof hk, cudaMemcpyHostToDevice);
  for (int t : {128, 256, 512}) { run<0>(t, dk, dout, sms); run<1>(t, dk, dout, sms); run<2>(t, dk, dout, sms); run<3>(t, dk, dout, sms); run<4>(t, dk, dout, sms); }
  return 0;
}
