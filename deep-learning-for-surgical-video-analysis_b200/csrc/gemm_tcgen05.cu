// Persistent, warp-specialised bf16 GEMM for sm_100a:  out = act(A * W^T + bias) (+ residual)
//
//   * operands staged global -> shared by TMA (cp.async.bulk.tensor.2d, 128B swizzle) into an mbarrier ring,
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=block_n) issued by ONE thread, fp32 accumulators in TMEM,
//   * two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * epilogue warps read TMEM with tcgen05.ld and fuse bias / GELU(erf) / ReLU / fp32 residual add / bf16 cast.
//
// Every nn.Linear, patchified nn.Conv2d and 1x1 conv of the LFB path goes through this kernel
// (reference: mix_transformer_evp.py:81-84 q/kv/proj, :37-40 fc1/fc2, :188 patch-embed conv, :89 sr conv,
// :599-642 adapter linears, :823-835 flow convs, :868 MHA projections; segformer_head.py:39,74 head).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (a warp may only touch TMEM lanes 32*(warp%4) .. +31, so 4 consecutive warps cover the 128 accumulator rows).
#include <stdio.h>

#include <algorithm>
#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace sv {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // 64 bf16 = 128 B = one swizzle row
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;                    // 2 accumulator buffers x 256 columns
constexpr int kAccStride = 256;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kSmemBudget = 196608;               // operand ring budget (bytes), + 1 KB alignment slack

__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.num_stages;
  const int b_tile_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = kATileBytes + b_tile_bytes;
  // 128B swizzle needs 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(&tmem_full_bar[a], 1);
        ptx::mbar_init(&tmem_empty_bar[a], 4);  // one arrive per epilogue warp
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&tmem_base_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int num_kb = (p.K + kBlockK - 1) / kBlockK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n_tiles) * kBlockM;
        const int n0 = (tile % p.num_n_tiles) * p.block_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kATileBytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
          ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m0);
          ptx::tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * kBlockK, n0);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16_f32(kBlockM, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kAccStride);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
          const uint64_t da = ptx::make_sw128_kmajor_desc(sa);
          const uint64_t db = ptx::make_sw128_kmajor_desc(sa + kATileBytes);
          const int k_left = p.K - kb * kBlockK;
          const int ksteps = k_left >= kBlockK ? kBlockK / 16 : (k_left + 15) / 16;
          for (int kk = 0; kk < ksteps; ++kk) {
            // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr>>4) start-address field
            ptx::umma_f16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);                       // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator ready for the epilogue
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_tiles) * kBlockM;
      const int n0 = (tile % p.num_n_tiles) * p.block_n;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + static_cast<uint32_t>(acc * kAccStride) + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        const int n = n0 + c0;
        if (n >= p.N) break;  // warp-uniform
        uint32_t r[16];
        ptx::tmem_ld_x16(t_row + static_cast<uint32_t>(c0), r);
        ptx::tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int ng = n + g * 8;
            if (ng < p.N) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ng));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ng + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (p.act != ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], p.act);
              }
              if (p.residual != nullptr) {
                const float* rp = p.residual + static_cast<long long>(row) * p.ldr + ng;
                const float4 r0 = *reinterpret_cast<const float4*>(rp);
                const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
                v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
              }
              if (p.out_fp32) {
                float* op = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldc + ng;
                *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
                bf16* op = reinterpret_cast<bf16*>(p.out) + static_cast<long long>(row) * p.ldc + ng;
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]);
                o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]);
                o.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(op) = o;
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, row stride ld elements; box = [box_rows, 64 cols], 128B swizzle, zero OOB fill.
int encode_operand_map(CUtensorMap* map, const bf16* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(SV_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(SV_ERR_INVALID, "GEMM operand base must be 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail(SV_ERR_INVALID, "GEMM operand row stride must be a multiple of 8 elements");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld ld=%lld box_rows=%d", static_cast<int>(r),
             static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld), box_rows);
    return fail(SV_ERR_CUDA, buf);
  }
  return SV_OK;
}

}  // namespace

int gemm_pick_block_n(int M, int N, int K, int num_sms) {
  // Candidates are legal UMMA N for M=128 (multiples of 16, <= 256).  Model: time ~ waves * (per-tile cost),
  // per-tile cost ~ fixed epilogue/pipeline overhead + N_tile * (k-blocks + epilogue share).
  const int n_pad = round_up(N, 16);
  const int m_tiles = ceil_div(M, kBlockM);
  const int kb = ceil_div(K, kBlockK);
  int best = 16;
  double best_cost = 1e30;
  for (int c = 256; c >= 16; c -= 16) {
    if (c > n_pad) continue;
    const int n_tiles = ceil_div(N, c);
    const long long tiles = static_cast<long long>(m_tiles) * n_tiles;
    const long long waves = (tiles + num_sms - 1) / num_sms;
    const double per_tile = 48.0 + c * (0.5 * kb + 1.0);  // MMA ~ c/2 "units" per k-block, epilogue ~ c units
    const double cost = waves * per_tile;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = c; }
  }
  return best;
}

int gemm_plan(const GemmDesc& d, GemmPlan* plan) {
  SV_CHECK(d.M > 0 && d.N > 0 && d.K > 0, "GEMM dims must be positive");
  SV_CHECK(d.K % 8 == 0 && d.N % 8 == 0, "GEMM needs K%8==0 and N%8==0");
  SV_CHECK(d.lda >= d.K && d.ldw >= d.K && d.ldc >= d.N, "GEMM leading dimensions too small");
  SV_CHECK(d.A && d.W && d.out, "GEMM null operand");
  SV_CHECK((reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ldc % (d.out_fp32 ? 4 : 8) == 0, "GEMM output must be 16B-aligned rows");
  if (d.residual) SV_CHECK((reinterpret_cast<uintptr_t>(d.residual) & 15) == 0 && d.ldr % 4 == 0 && d.ldr >= d.N, "GEMM residual alignment");
  if (d.bias) SV_CHECK((reinterpret_cast<uintptr_t>(d.bias) & 15) == 0, "GEMM bias alignment");
  const int sms = device_sm_count();
  SV_CHECK(sms > 0, "no CUDA device");
  GemmParams& p = plan->p;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.block_n = gemm_pick_block_n(d.M, d.N, d.K, sms);
  const int stage_bytes = kATileBytes + p.block_n * kBlockK * 2;
  p.num_stages = std::min(kMaxStages, kSmemBudget / stage_bytes);
  p.num_n_tiles = ceil_div(d.N, p.block_n);
  p.num_tiles = ceil_div(d.M, kBlockM) * p.num_n_tiles;
  p.act = d.act; p.out_fp32 = d.out_fp32;
  p.bias = d.bias; p.residual = d.residual; p.ldr = d.ldr; p.out = d.out; p.ldc = d.ldc;
  plan->grid = std::min(p.num_tiles, sms);
  plan->smem_bytes = static_cast<size_t>(p.num_stages) * stage_bytes + 1024;
  plan->flops = 2.0 * d.M * static_cast<double>(d.N) * d.K;
  SV_TRY(encode_operand_map(&plan->tmap_a, d.A, d.M, d.K, d.lda, kBlockM));
  SV_TRY(encode_operand_map(&plan->tmap_w, d.W, d.N, d.K, d.ldw, p.block_n));
  return SV_OK;
}

int gemm_launch(const GemmPlan& plan, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024);
  });
  if (attr_err != cudaSuccess) return fail(SV_ERR_CUDA, std::string("cudaFuncSetAttribute(gemm): ") + cudaGetErrorString(attr_err));
  gemm_bf16_tcgen05_kernel<<<plan.grid, kThreads, plan.smem_bytes, stream>>>(plan.tmap_a, plan.tmap_w, plan.p);
  return launch_status("gemm_bf16_tcgen05_kernel");
}

}  // namespace sv

extern "C" int sv_op_gemm_bf16(const uint16_t* A, int64_t lda, const uint16_t* W, int64_t ldw, int32_t M, int32_t N, int32_t K,
                               const float* bias, int32_t act, const float* residual, int64_t ldr, void* out, int64_t ldc,
                               int32_t out_fp32, void* stream) {
  sv::GemmDesc d;
  d.A = reinterpret_cast<const sv::bf16*>(A); d.lda = lda;
  d.W = reinterpret_cast<const sv::bf16*>(W); d.ldw = ldw;
  d.M = M; d.N = N; d.K = K; d.bias = bias; d.act = act; d.residual = residual; d.ldr = ldr;
  d.out = out; d.ldc = ldc; d.out_fp32 = out_fp32;
  sv::GemmPlan plan;
  SV_TRY(sv::gemm_plan(d, &plan));
  return sv::gemm_launch(plan, static_cast<cudaStream_t>(stream));
}
