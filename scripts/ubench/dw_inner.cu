// Micro-benchmark of the DWConv3x3+GELU "token pair" inner loop of mixffn.cu in isolation: the halo tile sits in shared memory, no TMA,
// no barriers.  Reports cycles per token pair per SM for 4/8/16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../deep-learning-for-surgical-video-analysis_b200/csrc -I ../../include -o dw_inner dw_inner.cu
#include <cstdio>
#include "common.cuh"
using namespace sv;
constexpr int W = 14, R = 7, ROWB = (W + 2) * 128;
template <int GELU>
__global__ void k(float* out, long long* cyc, int iters, int pairs_per_warp) {
  extern __shared__ uint8_t smem[];
  uint8_t* raw = smem;                       // (R+2) x (W+2) x 64 ch bf16
  uint8_t* sa = smem + (R + 2) * ROWB;       // 16 KB A tile
  for (int i = threadIdx.x; i < (R + 2) * ROWB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw)[i] = 0x3c003c00u + (i & 255);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  f32x2 wt[9];
  for (int t = 0; t < 9; ++t) wt[t] = f2_pack(0.1f * (t + 1) + lane * 1e-3f, 0.05f * t);
  const f32x2 bias2 = f2_pack(0.01f * lane, 0.02f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int j = 0; j < pairs_per_warp; ++j) {
      const int pj = (warp + j * 16) % 49;
      const int r0 = 2 * pj, ry = r0 / W, x = r0 - ry * W;
      const uint8_t* rp = raw + (ry * (W + 2) + x) * 128 + lane * 4;
      f32x2 a0 = bias2, a1 = bias2;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        f32x2 v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = f2_from_bf16x2(*reinterpret_cast<const uint32_t*>(rp + dy * ROWB + c * 128));
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          a0 = f2_fma(v[dx], wt[dy * 3 + dx], a0);
          a1 = f2_fma(v[dx + 1], wt[dy * 3 + dx], a1);
        }
      }
      if (GELU) f2_gelu_erf_poly_x2(a0, a1);
      float y0, y1, y2, y3;
      f2_unpack(a0, y0, y1);
      f2_unpack(a1, y2, y3);
      const int ao = r0 * 128 + (((lane >> 2) ^ (r0 & 7)) << 4) + (lane & 3) * 4;
      *reinterpret_cast<uint32_t*>(sa + ao) = pack_bf16x2(y0, y1);
      *reinterpret_cast<uint32_t*>(sa + ((ao + 128) ^ 16)) = pack_bf16x2(y2, y3);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = reinterpret_cast<float*>(sa)[threadIdx.x];
}
template <int GELU> void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int smem = (R + 2) * ROWB + 16384, iters = 400, ppw = 4;
  cudaFuncSetAttribute(k<GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) { k<GELU><<<148, warps * 32, smem>>>(out, cyc, iters, ppw); cudaDeviceSynchronize(); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("gelu=%d warps %2d: %.1f cycles per token pair per SM (%.0f cycles per pair per warp)  err=%s\n", GELU, warps, avg / (double(iters) * ppw * warps),
         avg / (double(iters) * ppw), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() { for (int w : {4, 8, 16, 24}) { run<1>(w); run<0>(w); } return 0; }
