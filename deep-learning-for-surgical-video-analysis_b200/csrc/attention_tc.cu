// softmax(Q K^T * scale) V on the 5th-generation tensor cores (tcgen05 + TMEM) for head_dim 64 and N_kv <= 448:
// every encoder block of mit_b1..b5_evp (mix_transformer_evp.py:123-127; N_kv = 49 at 224x224, 390/405 at 480x854).
//
// One CTA = one (frame, head) and a run of 128-row query tiles; K and V of the (frame, head) stay resident in shared memory.
//   warp 4   TMA: K/V tiles once, Q tiles double-buffered (3-D tensor maps [C, N, frames]: rows past the frame's last token are
//            zero-filled on load and clipped on store, so tiles never bleed into the next frame); O tiles back with a TMA store.
//   warp 5   one thread issues tcgen05.mma: S = Q K^T (M=128, N=64 per key tile, K = 64) into TMEM columns [0, 64*kt), then
//            O += P_t V_t per key tile (A = P from shared memory, B = V as an MN-major operand: V is used as stored, no transpose).
//   warps 0-3  softmax: thread = query row (TMEM lane).  tcgen05.ld of the S row, scale in the log2 domain, mask keys >= N_kv,
//            row max, exp2, row sum; P as bf16 into the 128B-swizzled K-major A tile (generic-proxy stores + fence.proxy.async);
//            after the PV MMAs: O row from TMEM, * 1/l, bf16, into the (consumed) Q buffer in the swizzled layout of the TMA store.
// With several key tiles (480x854) all of S (up to 448 columns) sits in TMEM at once: exact two-pass softmax, no online rescaling
// of O; the P tiles go through a 2-deep ring so the PV MMA of tile t overlaps the exponentials of tile t+1.
#include <stdlib.h>

#include <mutex>
#include <string>

#include "kernels.cuh"
#include "ptx.cuh"

namespace sv {
namespace {

constexpr int kHD = 64;                 // head dim == one 128-byte swizzle row of bf16
constexpr int kQRows = 128;             // query rows per tile == UMMA M == TMEM lanes
constexpr int kKeys = 64;               // keys per tile == UMMA N of S == K extent of one PV step group
constexpr int kMaxKt = 7;               // 7 * 64 S columns + 64 O columns = 512 TMEM columns
constexpr int kQBytes = kQRows * kHD * 2;   // 16 KB
constexpr int kKBytes = kKeys * kHD * 2;    // 8 KB
constexpr int kThreads = 192;

struct AttnTcParams {
  int Nq, Nkv, kt, tpc, qtiles;
  float scale_log2;
  unsigned tmem_cols;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :
               : "r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// MN-major B operand (V tile: rows = keys = K index, 128 B of head dim = N index per row), 128-byte swizzle, N = 64 = one atom:
// start>>4 | LBO (one N atom: unused) | SBO = 1024 B between 8-key groups | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// one 64-element bf16 row (8 chunks of 16 B) of a 128B-swizzled tile: chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
__device__ __forceinline__ void store_row_sw128(uint8_t* tile, int row, const uint32_t (&w)[32]) {
  uint8_t* rp = tile + row * 128;
  const int x = row & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(rp + ((c ^ x) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

template <bool MULTI>
__global__ void __launch_bounds__(kThreads, MULTI ? 1 : 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_o, const AttnTcParams p) {
  extern __shared__ uint8_t attn_tc_smem[];
  __shared__ __align__(8) uint64_t kv_full, q_full[2], s_full, p_full[2], p_empty[2], o_full, o_staged;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int tile0 = blockIdx.x * p.tpc;
  const int ntiles = min(p.tpc, p.qtiles - tile0);
  const int kt = MULTI ? p.kt : 1;
  uint8_t* smem = attn_tc_smem + ((1024u - (ptx::smem_u32(attn_tc_smem) & 1023u)) & 1023u);
  uint8_t* Qs = smem;                       // [2][16 KB]; tile j's buffer doubles as the staging tile of its output
  uint8_t* Ks = Qs + 2 * kQBytes;           // [kt][8 KB]
  uint8_t* Vs = Ks + kt * kKBytes;          // [kt][8 KB]
  uint8_t* Ps = Vs + kt * kKBytes;          // [MULTI ? 2 : 1][16 KB]

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_k);
    ptx::prefetch_tensormap(&tmap_v);
    ptx::prefetch_tensormap(&tmap_o);
  }
  if (warp == 5) {
    if (lane == 0) {
      ptx::mbar_init(&kv_full, 1);
      ptx::mbar_init(&q_full[0], 1);
      ptx::mbar_init(&q_full[1], 1);
      ptx::mbar_init(&s_full, 1);
      ptx::mbar_init(&p_full[0], 128);
      ptx::mbar_init(&p_full[1], 128);
      ptx::mbar_init(&p_empty[0], 1);
      ptx::mbar_init(&p_empty[1], 1);
      ptx::mbar_init(&o_full, 1);
      ptx::mbar_init(&o_staged, 128);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(&tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_o = tmem_base + static_cast<uint32_t>(kt * kKeys);
  const int c0 = head * kHD;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA: loads, and the stores of finished output tiles
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&kv_full, static_cast<uint32_t>(2 * kt * kKBytes));
      for (int t = 0; t < kt; ++t) {
        tma_load_3d(Ks + t * kKBytes, &tmap_k, &kv_full, c0, t * kKeys, b);
        tma_load_3d(Vs + t * kKBytes, &tmap_v, &kv_full, c0, t * kKeys, b);
      }
      for (int j = 0; j < 2 && j < ntiles; ++j) {
        ptx::mbar_arrive_expect_tx(&q_full[j], kQBytes);
        tma_load_3d(Qs + j * kQBytes, &tmap_q, &q_full[j], c0, (tile0 + j) * kQRows, b);
      }
      for (int j = 0; j < ntiles; ++j) {
        ptx::mbar_wait(&o_staged, static_cast<uint32_t>(j & 1));   // the 128 rows of O_j are staged in Q buffer j&1 (and fenced)
        tma_store_3d(&tmap_o, Qs + (j & 1) * kQBytes, c0, (tile0 + j) * kQRows, b);
        tma_store_commit();
        tma_store_wait_read();                                     // the store has read the buffer: it may be refilled / the CTA may exit
        if (j + 2 < ntiles) {
          ptx::mbar_arrive_expect_tx(&q_full[j & 1], kQBytes);
          tma_load_3d(Qs + (j & 1) * kQBytes, &tmap_q, &q_full[j & 1], c0, (tile0 + j + 2) * kQRows, b);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc_s = ptx::make_idesc_bf16_f32(kQRows, kKeys);                 // A, B K-major
      const uint32_t idesc_pv = ptx::make_idesc_bf16_f32(kQRows, kHD) | (1u << 16);     // B (= V) MN-major
      uint32_t ph_q[2] = {0u, 0u};
      uint32_t n_pf[2] = {0u, 0u};
      ptx::mbar_wait(&kv_full, 0u);
      for (int j = 0; j < ntiles; ++j) {
        const int qb = j & 1;
        ptx::mbar_wait(&q_full[qb], ph_q[qb]);
        ph_q[qb] ^= 1u;
        ptx::tc_fence_after();
        const uint64_t dq = ptx::make_sw128_kmajor_desc(ptx::smem_u32(Qs + qb * kQBytes));
        for (int t = 0; t < kt; ++t) {
          const uint64_t dk = ptx::make_sw128_kmajor_desc(ptx::smem_u32(Ks + t * kKBytes));
#pragma unroll
          for (int ks = 0; ks < kHD / 16; ++ks)
            ptx::umma_f16(tmem_base + static_cast<uint32_t>(t * kKeys), dq + static_cast<uint64_t>(ks * 2), dk + static_cast<uint64_t>(ks * 2), idesc_s, ks ? 1u : 0u);
        }
        ptx::umma_commit(&s_full);
        for (int t = 0; t < kt; ++t) {
          const int pb = t & 1;
          ptx::mbar_wait(&p_full[pb], n_pf[pb] & 1u);
          ++n_pf[pb];
          if (t == 0 && j > 0) ptx::mbar_wait(&o_staged, static_cast<uint32_t>((j - 1) & 1));   // O_{j-1} has left TMEM
          ptx::tc_fence_after();
          const uint64_t dp = ptx::make_sw128_kmajor_desc(ptx::smem_u32(Ps + pb * kQBytes));
          const uint64_t dv = make_sw128_mnmajor_desc(ptx::smem_u32(Vs + t * kKBytes));
#pragma unroll
          for (int ks = 0; ks < kKeys / 16; ++ks)   // 16 keys per step: +32 B along P's rows, +2 swizzle atoms (2048 B) down V
            ptx::umma_f16(tmem_o, dp + static_cast<uint64_t>(ks * 2), dv + static_cast<uint64_t>(ks * (2048 >> 4)), idesc_pv, (t | ks) ? 1u : 0u);
          ptx::umma_commit(&p_empty[pb]);
        }
        ptx::umma_commit(&o_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output rows (thread = query row = TMEM lane)
    const int row = warp * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t t_o = tmem_o + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t ph_s = 0u, ph_o = 0u;
    uint32_t n_pe[2] = {0u, 0u};
    for (int j = 0; j < ntiles; ++j) {
      const bool active = (tile0 + j) * kQRows + warp * 32 < p.Nq;   // warp-uniform: rows past the frame's last query do no math
      ptx::mbar_wait(&s_full, ph_s);
      ph_s ^= 1u;
      ptx::tc_fence_after();
      float m = -INFINITY, l = 0.f;
      if (MULTI && active) {
        for (int t = 0; t < kt; ++t) {
          uint32_t r[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            ptx::tmem_ld_x32(t_s + static_cast<uint32_t>(t * kKeys + h * 32), r);
            ptx::tmem_ld_wait();
            const int kbase = t * kKeys + h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float s = __uint_as_float(r[i]);
              m = fmaxf(m, (kbase + i < p.Nkv) ? s : -INFINITY);
            }
          }
        }
        m *= p.scale_log2;   // scale > 0: max commutes with it
      }
      for (int t = 0; t < kt; ++t) {
        const int pb = t & 1;
        uint8_t* Pt = Ps + pb * kQBytes;
        if (n_pe[pb] > 0u) ptx::mbar_wait(&p_empty[pb], (n_pe[pb] - 1u) & 1u);   // the PV MMA that read this P buffer last has retired
        ++n_pe[pb];
        if (active) {
          uint32_t r0[32], r1[32];
          ptx::tmem_ld_x32(t_s + static_cast<uint32_t>(t * kKeys), r0);
          ptx::tmem_ld_x32(t_s + static_cast<uint32_t>(t * kKeys + 32), r1);
          ptx::tmem_ld_wait();
          const int kbase = t * kKeys;
          if (!MULTI) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              m = fmaxf(m, (i < p.Nkv) ? __uint_as_float(r0[i]) : -INFINITY);
              m = fmaxf(m, (32 + i < p.Nkv) ? __uint_as_float(r1[i]) : -INFINITY);
            }
            m *= p.scale_log2;
          }
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = (kbase + i < p.Nkv) ? exp2f(fmaf(__uint_as_float(r0[i]), p.scale_log2, -m)) : 0.f;
            const float a1 = (kbase + i + 1 < p.Nkv) ? exp2f(fmaf(__uint_as_float(r0[i + 1]), p.scale_log2, -m)) : 0.f;
            const float b0 = (kbase + 32 + i < p.Nkv) ? exp2f(fmaf(__uint_as_float(r1[i]), p.scale_log2, -m)) : 0.f;
            const float b1 = (kbase + 32 + i + 1 < p.Nkv) ? exp2f(fmaf(__uint_as_float(r1[i + 1]), p.scale_log2, -m)) : 0.f;
            l += (a0 + a1) + (b0 + b1);
            w[i >> 1] = pack_bf16x2(a0, a1);
            w[16 + (i >> 1)] = pack_bf16x2(b0, b1);
          }
          store_row_sw128(Pt, row, w);
        }
        ptx::fence_proxy_async_smem();   // generic-proxy stores of P -> visible to the tensor core
        ptx::tc_fence_before();          // this thread's TMEM reads of S are complete (S may be overwritten once every row arrived)
        ptx::mbar_arrive(&p_full[pb]);
      }
      ptx::mbar_wait(&o_full, ph_o);
      ph_o ^= 1u;
      ptx::tc_fence_after();
      if (active) {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_x32(t_o, r0);
        ptx::tmem_ld_x32(t_o + 32u, r1);
        ptx::tmem_ld_wait();
        const float inv = 1.0f / l;
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          w[i >> 1] = pack_bf16x2(__uint_as_float(r0[i]) * inv, __uint_as_float(r0[i + 1]) * inv);
          w[16 + (i >> 1)] = pack_bf16x2(__uint_as_float(r1[i]) * inv, __uint_as_float(r1[i + 1]) * inv);
        }
        store_row_sw128(Qs + (j & 1) * kQBytes, row, w);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&o_staged);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// [frames, N, ld] bf16 (token-major, row stride ld) seen as a 3-D tensor {cols, N, frames}; box = 64 columns x rows x 1 frame, 128B swizzle
int encode_tokens_map(CUtensorMap* map, const bf16* base, int cols, int N, int B, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(SV_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(N) * static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kHD), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SV_ERR_CUDA, "cuTensorMapEncodeTiled(attention) failed, CUresult " + std::to_string(static_cast<int>(r)));
  return SV_OK;
}

}  // namespace

bool attention_tc_enabled() {
  static const bool on = [] { const char* e = getenv("SURGVID_ATTN_TC"); return !(e && atoi(e) == 0); }();
  return on;
}

bool attention_tc_supported(int hd, int Nkv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const void* q, const void* k, const void* v, const void* o) {
  if (hd != kHD || Nkv < 1 || Nkv > kMaxKt * kKeys) return false;
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return false;
  return ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15) == 0;
}

int attention_tc_plan(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* o, int64_t ldo, int B, int heads,
                      int Nq, int Nkv, float scale, AttnTcPlan* plan) {
  SV_CHECK(attention_tc_supported(kHD, Nkv, ldq, ldk, ldv, ldo, q, k, v, o), "attention_tc: unsupported shape / alignment");
  SV_CHECK(B > 0 && B <= 65535 && heads > 0 && heads <= 65535 && Nq > 0, "attention_tc dims");
  const int cols = heads * kHD;
  SV_CHECK(ldq >= cols && ldk >= cols && ldv >= cols && ldo >= cols, "attention_tc leading dims");
  SV_TRY(encode_tokens_map(&plan->tmap_q, q, cols, Nq, B, ldq, kQRows));
  SV_TRY(encode_tokens_map(&plan->tmap_k, k, cols, Nkv, B, ldk, kKeys));
  SV_TRY(encode_tokens_map(&plan->tmap_v, v, cols, Nkv, B, ldv, kKeys));
  SV_TRY(encode_tokens_map(&plan->tmap_o, o, cols, Nq, B, ldo, kQRows));
  plan->B = B; plan->heads = heads; plan->Nq = Nq; plan->Nkv = Nkv;
  plan->kt = ceil_div(Nkv, kKeys);
  plan->qtiles = ceil_div(Nq, kQRows);
  plan->scale_log2 = scale * 1.4426950408889634f;
  // query tiles per CTA: as many as still leave a few CTAs per SM-slot (K/V are loaded once per CTA)
  int tpc = 1;
  const long long want = 8LL * std::max(1, device_sm_count());
  while (tpc < 8 && tpc * 2 <= plan->qtiles && static_cast<long long>(ceil_div(plan->qtiles, tpc * 2)) * heads * B >= want) tpc *= 2;
  plan->tpc = tpc;
  const int need_cols = plan->kt * kKeys + kHD;
  plan->tmem_cols = need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512);
  plan->smem_bytes = 2 * kQBytes + 2 * plan->kt * kKBytes + (plan->kt > 1 ? 2 : 1) * kQBytes + 1024;
  return SV_OK;
}

int attention_tc_launch(const AttnTcPlan& plan, cudaStream_t st) {
  AttnTcParams p;
  p.Nq = plan.Nq; p.Nkv = plan.Nkv; p.kt = plan.kt; p.tpc = plan.tpc; p.qtiles = plan.qtiles; p.scale_log2 = plan.scale_log2;
  p.tmem_cols = static_cast<unsigned>(plan.tmem_cols);
  dim3 grid(ceil_div(plan.qtiles, plan.tpc), plan.heads, plan.B);
  if (plan.kt > 1) {
    SV_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_tc_kernel<true>), 2 * kQBytes + 2 * kMaxKt * kKBytes + 2 * kQBytes + 1024));
    attention_tc_kernel<true><<<grid, kThreads, plan.smem_bytes, st>>>(plan.tmap_q, plan.tmap_k, plan.tmap_v, plan.tmap_o, p);
  } else {
    SV_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_tc_kernel<false>), 3 * kQBytes + 2 * kKBytes + 1024));
    attention_tc_kernel<false><<<grid, kThreads, plan.smem_bytes, st>>>(plan.tmap_q, plan.tmap_k, plan.tmap_v, plan.tmap_o, p);
  }
  return launch_status("attention_tc_kernel");
}

}  // namespace sv
