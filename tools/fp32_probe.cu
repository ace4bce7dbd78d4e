/* tools/fp32_probe.cu -- diagnostics: FP32 issue rate of the receiver's inner-loop instruction patterns at the occupancy the
 * pipeline kernel runs at (one CTA of 14 warps per SM = 3-4 warps per SM sub-partition), against the rate at full occupancy.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/fp32_probe tools/fp32_probe.cu
 * Patterns:  0  v = v*a; v = v+b          (one register source + one uniform source; the bench's FMUL+FADD peak kernel)
 *            1  acc += h * (A - B)        (the Hilbert tap: FADD reg,reg / FMUL reg,reg / FADD reg,reg; 8 accumulators)
 *            2  DF1 biquad section chain   (the cascades: 5 FMUL + 4 FADD per section-sample, 4 sections skewed)
 *            3  v2 = fma2(v2, a2, b2)      (packed FFMA2 chains, 8 independent register pairs; counted as 2 lane-instructions each)
 *            4  the Hilbert tap in packed form: d = fma2(B,-1,A); p = fma2(h,d,-0); acc = fma2(p,1,acc), 4 accumulator pairs
 * Prints lane-instructions/s and the fraction of 128 lanes x SMs x clock. */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo_of(u64 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo + hi; }

template <int PAT>
__global__ void __launch_bounds__(1024) probe(float *out, int iters, float a, float b) {
  float acc[8], A[16], B[16];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = a + (float)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 16; i++) { A[i] = b * (float)(i + 1) + a; B[i] = a * (float)(i + 3) - b; }
  if (PAT == 0) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 12; u++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { acc[i] = __fmul_rn(acc[i], a); acc[i] = __fadd_rn(acc[i], b); }
      }
    }
  } else if (PAT == 1) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        const float hk = A[kk] ;
#pragma unroll
        for (int r = 0; r < 8; r++) acc[r] = __fadd_rn(acc[r], __fmul_rn(hk, __fadd_rn(A[(r - kk) & 15], -B[(r + kk + 1) & 15])));
      }
#pragma unroll
      for (int i = 0; i < 16; i++) { A[i] = __fadd_rn(A[i], b); B[i] = __fadd_rn(B[i], a); } /* loop-variant windows: nothing can be hoisted */
    }
  } else if (PAT == 3) {
    u64 v[8]; const u64 a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = pk(acc[i], acc[i] + 1.0f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 12; u++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fma2(v[i], a2, b2);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = lo_of(v[i]);
  } else if (PAT == 4) {
    u64 ac[4], RA[8], RB[8]; const u64 one = pk(1.0f, 1.0f), mone = pk(-1.0f, -1.0f), mzero = pk(-0.0f, -0.0f), a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll
    for (int i = 0; i < 4; i++) ac[i] = pk(acc[i], acc[i + 4]);
#pragma unroll
    for (int i = 0; i < 8; i++) { RA[i] = pk(A[i], A[i + 8]); RB[i] = pk(B[i], B[i + 8]); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        const u64 hk = RA[kk];
#pragma unroll
        for (int r = 0; r < 4; r++) ac[r] = fma2(fma2(hk, fma2(RB[(r + kk + 1) & 7], mone, RA[(r - kk) & 7]), mzero), one, ac[r]);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) { RA[i] = fma2(RA[i], one, b2); RB[i] = fma2(RB[i], one, a2); }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = lo_of(ac[i]);
  } else {
    float c[20], s[16];
#pragma unroll
    for (int i = 0; i < 20; i++) c[i] = a * (float)(i + 1) * 0.01f;
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = b * (float)i;
    float p0 = a, p1 = b, p2 = a + b, v = a;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float in[4] = {v, p0, p1, p2}, o[4];
#pragma unroll
        for (int k = 3; k >= 0; k--) {
          float t = __fmul_rn(c[5 * k], in[k]);
          t = __fadd_rn(t, __fmul_rn(c[5 * k + 1], s[4 * k]));
          t = __fadd_rn(t, __fmul_rn(c[5 * k + 2], s[4 * k + 1]));
          t = __fadd_rn(t, __fmul_rn(c[5 * k + 3], s[4 * k + 2]));
          t = __fadd_rn(t, __fmul_rn(c[5 * k + 4], s[4 * k + 3]));
          s[4 * k + 1] = s[4 * k]; s[4 * k] = in[k]; s[4 * k + 3] = s[4 * k + 2]; s[4 * k + 2] = t;
          o[k] = t;
        }
        p0 = o[0]; p1 = o[1]; p2 = o[2]; v = o[3] * 1e-3f + a;
      }
    }
    acc[0] += p0 + p1 + p2 + v;
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) sum += acc[i];
  if (sum == 123.456f) out[0] = sum;
}

static double run(int pat, int blocks_per_sm, int threads, int iters, int sms) {
  float *d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    if (pat == 0) probe<0><<<sms * blocks_per_sm, threads>>>(d, iters, 0.999f, 0.001f);
    else if (pat == 1) probe<1><<<sms * blocks_per_sm, threads>>>(d, iters, 0.999f, 0.001f);
    else if (pat == 2) probe<2><<<sms * blocks_per_sm, threads>>>(d, iters, 0.999f, 0.001f);
    else if (pat == 3) probe<3><<<sms * blocks_per_sm, threads>>>(d, iters, 0.999f, 0.001f);
    else probe<4><<<sms * blocks_per_sm, threads>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const double per_iter = pat == 0 ? 12 * 16.0 : (pat == 1 ? 8 * 24.0 + 32 : (pat == 2 ? 4 * 36.0 + 4 * 2 : (pat == 3 ? 12 * 8 * 2.0 : (8 * 12 + 16) * 2.0)));
  cudaFree(d);
  return (double)sms * blocks_per_sm * threads * (double)iters * per_iter / (ms * 1e-3);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const double peak = (double)sms * 128.0 * khz * 1e3;
  printf("SMs %d, clock %.0f MHz, 128-lane peak %.2f T lane-instr/s\n", sms, khz / 1e3, peak / 1e12);
  const int cfg[][2] = {{1, 128}, {1, 256}, {1, 448}, {1, 512}, {1, 1024}};
  for (int pat = 0; pat < 5; pat++)
    for (auto &c : cfg) {
      const double r = run(pat, c[0], c[1], 20000, sms);
      printf("pattern %d  %d x %4d threads/SM (%4.1f warps per sub-partition): %7.2f T lane-instr/s = %5.1f %% of peak\n", pat, c[0], c[1],
             c[0] * c[1] / 128.0, r / 1e12, 100.0 * r / peak);
    }
  return 0;
}
