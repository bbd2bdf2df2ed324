// Shared host/device helpers for libsurgvid (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/surgvid.h"

namespace sv {

// ---- error plumbing: integer status across the C ABI, message via sv_last_error() ----------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define SV_CUDA_OK(expr)                                                                           \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::sv::fail(SV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

#define SV_CHECK(cond, msg)                                                                        \
  do {                                                                                             \
    if (!(cond)) return ::sv::fail(SV_ERR_INVALID, std::string(msg) + " [" #cond "]");             \
  } while (0)

#define SV_TRY(expr)                                                                               \
  do {                                                                                             \
    int _s = (expr);                                                                               \
    if (_s != SV_OK) return _s;                                                                    \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SV_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return SV_OK;
}

int device_sm_count();

typedef __nv_bfloat16 bf16;

// activation codes shared by GEMM epilogue and elementwise kernels
enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

#ifdef __CUDACC__
// exact (erf) GELU, as nn.GELU() default — mix_transformer_evp.py:33,39
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU(erf) with erf from Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 rounding of the result):
// one MUFU.EX2 + one MUFU.RCP instead of erff's branchy polynomial; used where the activation output is stored as bf16.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);  // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace sv
