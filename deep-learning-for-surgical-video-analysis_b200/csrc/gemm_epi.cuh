// Epilogue building blocks shared by the tcgen05 kernels (gemm_tcgen05.cu, mixffn.cu): accumulator rows come out of TMEM with
// thread = row; a per-warp shared-memory staging tile transposes 32x32 fp32 blocks so that bias / activation / fp32 residual /
// cast and the global loads+stores run with 8 lanes on one 128-byte row segment.
#pragma once
#include "common.cuh"

namespace sv {

constexpr int kStageLd = 36;  // epilogue staging row stride (floats): 32 columns + 4 pad (conflict-free)

// residual rows for one 32-column chunk: 8 passes x (4 rows x 8 lanes x float4)
template <bool RESID>
__device__ __forceinline__ void epi_load_residual(float4 (&res)[8], const float* residual, long long ldr, int row_limit, int row_base, int n,
                                                  bool col_ok) {
  if constexpr (RESID) {
#pragma unroll
    for (int ps = 0; ps < 8; ++ps) {
      const int row = row_base + ps * 4;
      res[ps] = (col_ok && row < row_limit) ? *reinterpret_cast<const float4*>(residual + static_cast<long long>(row) * ldr + n)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// accumulator registers (thread = row) -> staging tile (transpose point)
__device__ __forceinline__ void epi_park(float* stg, int lane, const uint32_t (&r)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * kStageLd + 4 * j) =
        make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
}

// staging tile -> bias / activation / residual / cast -> global, 8 lanes per 128-byte row segment
template <int ACT, bool OUT_F32, bool RESID>
__device__ __forceinline__ void epi_store(void* out, long long ldc, int row_limit, const float* stg, const float* bias_s, const float4 (&res)[8],
                                          int row_base, int n, int c_local, bool col_ok, int sub_row, int c4) {
  const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c_local);
  const long long out_off = static_cast<long long>(row_base) * ldc + n;
  const long long row_step = 4 * ldc;
#pragma unroll
  for (int ps = 0; ps < 8; ++ps) {
    const int row = row_base + ps * 4;
    float4 v = *reinterpret_cast<const float4*>(stg + (ps * 4 + sub_row) * kStageLd + c4);
    v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
    if constexpr (ACT == ACT_GELU) {
      if constexpr (OUT_F32) {
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
      } else {
        // bf16 result: the packed degree-8 erf polynomial (|error| <= 6e-5, below the bf16 rounding) — two FFMA2 chains instead of
        // four scalar exp-based evaluations; the GELU GEMMs of the adapter have K = 16..128 and are bound by this epilogue
        f32x2 lo = f2_pack(v.x, v.y), hi = f2_pack(v.z, v.w);
        f2_gelu_erf_poly_x2(lo, hi);
        f2_unpack(lo, v.x, v.y);
        f2_unpack(hi, v.z, v.w);
      }
    } else if constexpr (ACT == ACT_RELU) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if constexpr (RESID) { v.x += res[ps].x; v.y += res[ps].y; v.z += res[ps].z; v.w += res[ps].w; }
    if (col_ok && row < row_limit) {
      if constexpr (OUT_F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + out_off + ps * row_step) = v;
      } else {
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out) + out_off + ps * row_step) = o;
      }
    }
  }
}

}  // namespace sv
