// Micro-benchmark: issue cost per warp instruction and per scheduler of FFMA, FFMA2 (register and immediate forms) and FMUL2 on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_tput fma_tput.cu     Run: ./fma_tput
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int ITERS = 2048, CH = 8;

template <int MODE>
__global__ void k(float* out, long long* cyc, float x0) {
  float a[CH]; f32x2 v[CH];
  for (int i = 0; i < CH; ++i) { a[i] = x0 + i + threadIdx.x; v[i] = pk(a[i], a[i] + 0.5f); }
  const float m = 0.999f + x0 * threadIdx.x; const float cc = x0 * (threadIdx.x + 3); const f32x2 m2 = pk(m, m + x0), c2 = pk(cc, cc + 1e-3f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) a[i] = fma1(a[i], m, cc);                                  // FFMA, 3 registers
      if (MODE == 1) v[i] = fma2(v[i], m2, c2);                                  // FFMA2, 3 register pairs
      if (MODE == 2) v[i] = fma2(v[i], m2, pk(1.25e-3f, 1.25e-3f));              // FFMA2, immediate addend
      if (MODE == 3) v[i] = mul2(v[i], m2);                                      // FMUL2
      if (MODE == 4) a[i] = fma1(a[i], m, 1.25e-3f);                             // FFMA, immediate addend
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < CH; ++i) { s += a[i]; float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, threads>>>(out, cyc, 1e-6f); cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(out, cyc, 1e-6f); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double warps_per_smsp = threads / 32.0 / 4.0;
  const double inst_per_smsp = warps_per_smsp * ITERS * CH;
  printf("%-28s threads %4d: %8.0f cycles  -> %.2f cycles per warp-instruction per scheduler (%.1f lane-FMA/clk/SM)\n", name, threads, avg, avg / inst_per_smsp,
         (MODE == 0 || MODE == 4 ? 32.0 : 64.0) * 4.0 / (avg / inst_per_smsp));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 512, 1024}) {
    run<0>("FFMA  reg,reg,reg", th); run<4>("FFMA  reg,reg,imm", th); run<1>("FFMA2 reg,reg,reg", th); run<2>("FFMA2 reg,reg,imm", th); run<3>("FMUL2 reg,reg", th);
  }
  return 0;
}
