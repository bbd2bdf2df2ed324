// Micro-benchmark: what one 32-column chunk of the GEMM's bf16 TMA epilogue costs on sm_100a, piece by piece.
// 8 epilogue warps (two per TMEM lane quarter) each run `iters` chunks: tcgen05.ld 32x32b.x32 (+wait) -> bias add + bf16 pack ->
// 4 x STS.128 (64B-swizzled staging) -> fence.proxy.async -> __syncwarp -> [TMA store + commit + wait_group.read].
// FLAGS: 1 tcgen05.ld, 2 math + STS, 4 fence.proxy.async, 8 TMA store (to a scratch tensor), 16 prefetch the next chunk's tcgen05.ld
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../deep-learning-for-surgical-video-analysis_b200/csrc -o epi_rate epi_rate.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace sv;

__device__ __forceinline__ uint32_t pack2(float a, float b) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b)); return r; }

template <int FLAGS, int BUFS>
__global__ void __launch_bounds__(320, 1) k(long long* cyc, int iters, const __grid_constant__ CUtensorMap tmap, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[8][256];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&bias_s[0][0])[i] = 0.001f * i;
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp >= 2) {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const uint32_t t_row = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* sbase = smem + (warp - 2) * 4 * 2048;
    constexpr uint32_t NB = BUFS;
    const float* bs = bias_s[warp - 2];
    uint32_t r[32], r2[32];
    for (int i = 0; i < 32; ++i) { r[i] = 0x3f800000u + lane + i; r2[i] = r[i]; }
    float keep = 0.f;
    uint32_t n_store = 0;
    const int x = (lane >> 1) & 3;
    __syncwarp();
    const long long t0 = clock64();
    if (FLAGS & 16) ptx::tmem_ld_x32(t_row + half * 32, r);
    for (int it = 0; it < iters; ++it) {
      const int c = (half + 2 * it) & 7;
      if (FLAGS & 1) {
        if (FLAGS & 16) { ptx::tmem_ld_wait(); for (int i = 0; i < 32; ++i) r2[i] = r[i]; ptx::tmem_ld_x32(t_row + ((c + 2) & 7) * 32, r); }
        else { ptx::tmem_ld_x32(t_row + c * 32, r2); ptx::tmem_ld_wait(); }
      }
      uint8_t* sb = sbase + (n_store % NB) * 2048;
      if ((FLAGS & 8) && n_store >= NB) { if (ptx::elect_one()) ptx::bulk_wait_group_read(NB - 1); __syncwarp(); }
      if (FLAGS & 2) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bs + c * 32 + i);
          w[i >> 1] = pack2(__uint_as_float(r2[i]) + b4.x, __uint_as_float(r2[i + 1]) + b4.y);
          w[(i >> 1) + 1] = pack2(__uint_as_float(r2[i + 2]) + b4.z, __uint_as_float(r2[i + 3]) + b4.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(sb + lane * 64 + ((j ^ x) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      } else {
        keep += __uint_as_float(r2[it & 31]);
      }
      if (FLAGS & 4) ptx::fence_proxy_async_smem();
      __syncwarp();
      if (FLAGS & 8) {
        if (ptx::elect_one()) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmap)),
                       "r"(ptx::smem_u32(sb)), "r"(c * 32), "r"(static_cast<int>(blockIdx.x) * 128 + quarter * 32) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      ++n_store;
    }
    if (FLAGS & 8) { if (ptx::elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); __syncwarp(); }
    if (FLAGS & 1) ptx::tmem_ld_wait();
    const long long t1 = clock64();
    if (lane == 0 && warp == 2) cyc[blockIdx.x] = t1 - t0;
    if (keep == 123.f) sink[threadIdx.x] = keep + __uint_as_float(r[3]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}


// 64-column variant: two tcgen05.ld x32 -> 8 STS.128 per row in the 128B-swizzled layout -> ONE TMA store of a 64-column x 32-row box (4 KB)
template <int BUFS>
__global__ void __launch_bounds__(320, 1) kwide(long long* cyc, int iters, const __grid_constant__ CUtensorMap tmap, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[8][256];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&bias_s[0][0])[i] = 0.001f * i;
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp >= 2) {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const uint32_t t_row = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* sbase = smem + (warp - 2) * BUFS * 4096;
    const float* bs = bias_s[warp - 2];
    uint32_t ra[32], rb[32];
    uint32_t n_store = 0;
    const int x = lane & 7;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {   // one iteration = 64 columns
      const int c = (half + 2 * it) & 3;   // 64-column chunk index
      ptx::tmem_ld_x32(t_row + c * 64, ra);
      ptx::tmem_ld_x32(t_row + c * 64 + 32, rb);
      uint8_t* sb = sbase + (n_store % BUFS) * 4096;
      if (n_store >= BUFS) { if (ptx::elect_one()) ptx::bulk_wait_group_read(BUFS - 1); __syncwarp(); }
      ptx::tmem_ld_wait();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t* r = h ? rb : ra;
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bs + c * 64 + h * 32 + i);
          w[i >> 1] = pack2(__uint_as_float(r[i]) + b4.x, __uint_as_float(r[i + 1]) + b4.y);
          w[(i >> 1) + 1] = pack2(__uint_as_float(r[i + 2]) + b4.z, __uint_as_float(r[i + 3]) + b4.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(sb + lane * 128 + (((h * 4 + j) ^ x) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (ptx::elect_one()) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmap)),
                     "r"(ptx::smem_u32(sb)), "r"(c * 64), "r"(static_cast<int>(blockIdx.x) * 128 + quarter * 32) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      ++n_store;
    }
    if (ptx::elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    const long long t1 = clock64();
    if (lane == 0 && warp == 2) cyc[blockIdx.x] = t1 - t0;
    if (ra[0] == 123u) sink[threadIdx.x] = 1.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int FLAGS, int BUFS = 4> void run(const char* name, const CUtensorMap& tmap) {
  long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4096);
  const int iters = 512, smem = 8 * 4 * 2048 + 2048;
  cudaFuncSetAttribute(k<FLAGS, BUFS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) { k<FLAGS, BUFS><<<148, 320, smem>>>(cyc, iters, tmap, sink); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; } }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-64s %7.1f cycles per chunk per warp -> a 128x256 tile (4 chunks per warp) takes %6.0f cycles\n", name, avg / iters, 4 * avg / iters);
  cudaFree(cyc); cudaFree(sink);
}
template <int BUFS> void runwide(const char* name, const CUtensorMap& tmap) {
  long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4096);
  const int iters = 256, smem = 8 * BUFS * 4096 + 2048;
  cudaFuncSetAttribute(kwide<BUFS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) { kwide<BUFS><<<148, 320, smem>>>(cyc, iters, tmap, sink); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; } }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-64s %7.1f cycles per 32-column chunk per warp -> a 128x256 tile takes %6.0f cycles\n", name, avg / iters / 2, 2 * avg / iters);
  cudaFree(cyc); cudaFree(sink);
}
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fnp);
  void* out; cudaMalloc(&out, 148ull * 128 * 256 * 2);
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {256, 148 * 128}; cuuint64_t gstr[1] = {512}; cuuint32_t box[2] = {32, 32}; cuuint32_t estr[2] = {1, 1};
  fn(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  run<1>("tcgen05.ld x32 + wait", tmap);
  run<1 | 16>("tcgen05.ld x32, next chunk prefetched", tmap);
  run<2>("bias add + pack + 4 STS.128", tmap);
  run<2 | 4>("bias add + pack + STS + fence.proxy.async", tmap);
  run<1 | 2>("ld + math + STS", tmap);
  run<1 | 2 | 4>("ld + math + STS + fence", tmap);
  run<1 | 2 | 4 | 16>("ld(prefetched) + math + STS + fence", tmap);
  run<1 | 2 | 4 | 8>("ld + math + STS + fence + TMA store", tmap);
  run<1 | 2 | 4 | 8 | 16>("ld(prefetched) + math + STS + fence + TMA store  (= the kernel)", tmap);
  run<2 | 4 | 8>("math + STS + fence + TMA store", tmap);
  run<1 | 2 | 4 | 8 | 16, 2>("the kernel with 2 staging buffers per warp", tmap);
  CUtensorMap tmapw;
  cuuint32_t boxw[2] = {64, 32};
  fn(&tmapw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, boxw, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  runwide<2>("64-column boxes (128 B rows, SW128), 2 x 4 KB buffers per warp", tmapw);
  runwide<3>("64-column boxes (128 B rows, SW128), 3 x 4 KB buffers per warp", tmapw);
  return 0;
}
