// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS operands, 128B swizzle) on sm_100a
//   mode 0  back to back on one accumulator, one commit at the end                      -> the tensor pipe's own rate
//   mode 1  groups of 4 + commit per group, issuer waits for group g-S before group g     -> commit / mbarrier recycle cost
//   mode 2  as 1, but the recycled slot is handed back through a second ("producer") thread as in gemm_tcgen05.cu (no loads)
//   mode 4  commit per group to a barrier nobody waits on                                -> cost of tcgen05.commit itself
//   mode 3  as 0 with four warps reading the OTHER accumulator with tcgen05.ld at full speed -> TMEM read interference
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../deep-learning-for-surgical-video-analysis_b200/csrc -o umma_rate umma_rate.cu -lcuda
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace sv;

constexpr int kStages = 8;
constexpr int kStageBytes = 16384 + 32768;

__device__ __forceinline__ void spin_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
template <int MODE, int G>
__global__ void __launch_bounds__(192, 1) k(long long* cyc, int N, int groups, int S, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < S * kStageBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 255);  // finite bf16 values
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_barrier_init();
    stop = 0;
  }
  ptx::fence_proxy_async_smem();
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, N);
  if (warp == 0 && (MODE == 8 || MODE == 9)) {
    // warp-converged issuer: every lane runs the loop and polls the barrier, one elected lane issues the MMAs and the commit
    const long long t0 = clock64();
    int stage = 0; uint32_t phase = 0;
    for (int g = 0; g < groups; ++g) {
      if (MODE == 8 && g >= S) ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (MODE == 9) { ptx::mbar_wait(&full_bar[stage], phase); ptx::tc_fence_after(); }
      const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
      const uint64_t da = ptx::make_sw128_kmajor_desc(sa), db = ptx::make_sw128_kmajor_desc(sa + 16384);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < G; ++kk) ptx::umma_f16(tmem, da + (kk & 3) * 2, db + (kk & 3) * 2, idesc, (g | kk) != 0);
        ptx::umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) ptx::umma_commit(&done_bar);
    __syncwarp();
    ptx::mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    if (lane == 0) { cyc[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp == 0 && lane == 0) {
    const long long t0 = clock64();
    int stage = 0; uint32_t phase = 0;
    for (int g = 0; g < groups; ++g) {
      if (MODE == 1 && g >= S) ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (MODE == 7 && g >= S) spin_test_wait(&empty_bar[stage], phase ^ 1u);   // (phase bookkeeping: the slot was committed one pass ago)
      if (MODE == 2) { ptx::mbar_wait(&full_bar[stage], phase); ptx::tc_fence_after(); }
      const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
      const uint64_t da = ptx::make_sw128_kmajor_desc(sa), db = ptx::make_sw128_kmajor_desc(sa + 16384);
#pragma unroll
      for (int kk = 0; kk < G; ++kk) ptx::umma_f16(tmem, da + (kk & 3) * 2, db + (kk & 3) * 2, idesc, (g | kk) != 0);
      if (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6 || MODE == 7) ptx::umma_commit(&empty_bar[stage]);
      if (MODE == 5) ptx::mbar_wait(&empty_bar[stage], phase);        // serial: wait for this very group (try_wait)
      if (MODE == 6) spin_test_wait(&empty_bar[stage], phase);        // serial, polling with test_wait
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    ptx::umma_commit(&done_bar);
    ptx::mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (warp == 1 && lane == 0 && (MODE == 2 || MODE == 9)) {
    int stage = 0; uint32_t phase = 0;
    for (int g = 0; g < groups; ++g) {
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
      ptx::mbar_arrive(&full_bar[stage]);
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= 2 && MODE == 3) {
    const uint32_t t_row = tmem + 256u + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t r[32]; float acc = 0.f;
    while (!stop) {
      for (int c = 0; c < 8; ++c) { ptx::tmem_ld_x32(t_row + c * 32, r); ptx::tmem_ld_wait(); acc += __uint_as_float(r[lane & 31]); }
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int MODE, int G> void run(const char* name, int N, int S) {
  long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4096);
  const int groups = 4096 / G, smem = S * kStageBytes + 2048;
  cudaFuncSetAttribute(k<MODE, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
  for (int rep = 0; rep < 2; ++rep) { k<MODE, G><<<148, 192, smem>>>(cyc, N, groups, S, sink); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; } }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-44s N=%3d S=%d MMAs/commit=%2d: %7.1f cycles per MMA (floor N/2 = %d), %7.1f per commit group\n", name, N, S, G, avg / (groups * (double)G), N / 2, avg / groups);
  cudaFree(cyc); cudaFree(sink);
}
template <int G> void sweep(int N) {
  run<4, G>("commit per group, nobody waits", N, 4);
  run<1, G>("commit per group, issuer recycles slot", N, 4);
  run<2, G>("commit per group, slot via producer thread", N, 4);
  run<8, G>("converged warp + elect, issuer recycles", N, 4);
  run<9, G>("converged warp + elect, producer thread", N, 4);
}
int main() {
  for (int N : {256, 160, 64}) {
    run<0, 4>("back to back", N, 4);
    run<3, 4>("back to back + 4 warps tcgen05.ld", N, 4);
    sweep<1>(N); sweep<4>(N); sweep<8>(N);
  }
  return 0;
}
