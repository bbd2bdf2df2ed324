// Trans-SVNet head: the inner module of adapter_transformer.Transformer (adapter_transformer.py:317-325, 348),
//   output[t] = Transformer2_3_1(inputs[t] = the len_q-frame causal window of the MS-TCN logits ending at t, feas[t] = tanh(fc(LFB[t])))
// fused into ONE kernel that reads the channel-major MS-TCN logits and the query directly: the [T, len_q, 14] window tensor the
// reference builds with a Python loop is never materialised.
//
// The source of Transformer2_3_1 is NOT in the reference tree; the arithmetic here follows the published upstream
// architecture as restated in oracle/trans_head_oracle.py (PARITY UNPINNED): one encoder layer over the window (4-head self-attention,
// residual + LayerNorm, position-wise FFN with ReLU, residual + LayerNorm), one decoder layer whose single query cross-attends the
// encoder output (same blocks).  fp32 throughout (d_model = 14: nothing here is tensor-core shaped).
//
// One persistent CTA = two groups of 128 threads (4 warps = 4 heads each); every group walks its own frames t with its own Q/K/V tiles
// and named barriers, and both share the weights (86 KB, d_model padded to 16) in shared memory.
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace sv {
namespace {

constexpr int kDP = 16;      // d_model padded
constexpr int kH = 4;        // heads
constexpr int kDK = 32;      // d_k = d_v
constexpr int kHD = kH * kDK;
constexpr int kLMax = 32;    // len_q <= 32
constexpr int kFFMax = 64;   // d_ff <= 64
constexpr int kLdS = kHD + 1;  // row stride of the Q/K/V tiles (conflict-free row-wise reads)

// blob layout of one attention block / one FFN block (floats)
constexpr int kAttnWQ = 0, kAttnBQ = kHD * kDP, kAttnWK = kAttnBQ + kHD, kAttnBK = kAttnWK + kHD * kDP, kAttnWV = kAttnBK + kHD,
              kAttnBV = kAttnWV + kHD * kDP, kAttnWO = kAttnBV + kHD, kAttnBO = kAttnWO + kDP * kHD, kAttnG = kAttnBO + kDP, kAttnB = kAttnG + kDP,
              kAttnSize = kAttnB + kDP;
constexpr int kFfnW1 = 0, kFfnB1 = kFFMax * kDP, kFfnW2 = kFfnB1 + kFFMax, kFfnB2 = kFfnW2 + kDP * kFFMax, kFfnG = kFfnB2 + kDP, kFfnB = kFfnG + kDP,
              kFfnSize = kFfnB + kDP;
constexpr int kBlobSize = 2 * (kAttnSize + kFfnSize);   // encoder attn | encoder ffn | decoder attn | decoder ffn

struct TransParams {
  int D, L, FF, n_videos;
  float scale;   // 1 / sqrt(d_k)
};

__device__ __forceinline__ void layer_norm_inplace(float (&v)[kDP], int D, const float* g, const float* b) {
  float m = 0.f;
#pragma unroll
  for (int c = 0; c < kDP; ++c) m += (c < D) ? v[c] : 0.f;
  m /= static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kDP; ++c) { const float d = (c < D) ? v[c] - m : 0.f; q += d * d; }
  const float r = rsqrtf(q / static_cast<float>(D) + 1e-5f);
#pragma unroll
  for (int c = 0; c < kDP; ++c) v[c] = (c < D) ? (v[c] - m) * r * g[c] + b[c] : 0.f;
}

constexpr int kGroups = 2;                                       // 128-thread groups per CTA
constexpr int kGroupFloats = 3 * kLMax * kLdS + 3 * kLMax * kDP + kDP;   // per-group tiles
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

__global__ void __launch_bounds__(128 * kGroups) trans_head_kernel(const float* __restrict__ logits, int64_t ldx, const float* __restrict__ query,
                                                         const int64_t* __restrict__ offsets, int64_t T, const float* __restrict__ blob,
                                                         float* __restrict__ out, const TransParams p) {
  extern __shared__ __align__(16) float sm[];
  float* Wb = sm;                              // kBlobSize
  const int grp = threadIdx.x >> 7;
  float* Qs = Wb + kBlobSize + grp * kGroupFloats;   // [kLMax][kLdS]  Q, then the attention context
  float* Ks = Qs + kLMax * kLdS;
  float* Vs = Ks + kLMax * kLdS;
  float* Xs = Vs + kLMax * kLdS;               // [kLMax][kDP]  window
  float* Ys = Xs + kLMax * kDP;                // [kLMax][kDP]
  float* Es = Ys + kLMax * kDP;                // [kLMax][kDP]  encoder output
  float* qv = Es + kLMax * kDP;                // [kDP] decoder query
  const int tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;
  const int D = p.D, L = p.L, FF = p.FF;
  for (int i = threadIdx.x; i < kBlobSize / 4; i += 128 * kGroups) reinterpret_cast<float4*>(Wb)[i] = __ldg(reinterpret_cast<const float4*>(blob) + i);
  const float* EA = Wb;
  const float* EF = EA + kAttnSize;
  const float* DA = EF + kFfnSize;
  const float* DF = DA + kAttnSize;
  __syncthreads();

  for (int64_t t = static_cast<int64_t>(blockIdx.x) * kGroups + grp; t < T; t += static_cast<int64_t>(gridDim.x) * kGroups) {
    // first frame of t's video (zero left padding of the window never crosses a video boundary, adapter_transformer.py:336-341)
    int lo = 0, hi = p.n_videos;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (offsets[mid] <= t) lo = mid; else hi = mid;
    }
    const int64_t start = offsets[lo];
    // ---- (1) window X [L, D] and the decoder query
    for (int idx = tid; idx < L * kDP; idx += 128) {
      const int i = idx / kDP, c = idx % kDP;
      const int64_t src = t - (L - 1) + i;
      Xs[idx] = (c < D && src >= start) ? __ldg(logits + static_cast<int64_t>(c) * ldx + src) : 0.f;
    }
    if (tid < kDP) qv[tid] = tid < D ? __ldg(query + t * D + tid) : 0.f;
    group_sync(grp);
    // ---- (2) encoder Q, K, V: thread = one of the 128 projection columns
    {
      float wq[kDP], wk[kDP], wv[kDP];
#pragma unroll
      for (int c = 0; c < kDP; ++c) { wq[c] = EA[kAttnWQ + tid * kDP + c]; wk[c] = EA[kAttnWK + tid * kDP + c]; wv[c] = EA[kAttnWV + tid * kDP + c]; }
      const float bq = EA[kAttnBQ + tid], bk = EA[kAttnBK + tid], bv = EA[kAttnBV + tid];
      for (int i = 0; i < L; ++i) {
        float q = bq, k = bk, v = bv;
#pragma unroll
        for (int c = 0; c < kDP; ++c) { const float x = Xs[i * kDP + c]; q = fmaf(x, wq[c], q); k = fmaf(x, wk[c], k); v = fmaf(x, wv[c], v); }
        Qs[i * kLdS + tid] = q; Ks[i * kLdS + tid] = k; Vs[i * kLdS + tid] = v;
      }
    }
    group_sync(grp);
    // ---- (3) self-attention: warp = head, lane = query row
    if (lane < L) {
      float q[kDK];
#pragma unroll
      for (int d = 0; d < kDK; ++d) q[d] = Qs[lane * kLdS + warp * kDK + d];
      float s[kLMax];
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < kLMax; ++j) {
        float a = 0.f;
        if (j < L) {
#pragma unroll
          for (int d = 0; d < kDK; ++d) a = fmaf(q[d], Ks[j * kLdS + warp * kDK + d], a);
          a *= p.scale;
          m = fmaxf(m, a);
        }
        s[j] = a;
      }
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < kLMax; ++j) { s[j] = (j < L) ? __expf(s[j] - m) : 0.f; den += s[j]; }
      const float inv = 1.0f / den;
      float ctx[kDK];
#pragma unroll
      for (int d = 0; d < kDK; ++d) ctx[d] = 0.f;
#pragma unroll
      for (int j = 0; j < kLMax; ++j) {
        if (j < L) {
          const float pj = s[j] * inv;
#pragma unroll
          for (int d = 0; d < kDK; ++d) ctx[d] = fmaf(pj, Vs[j * kLdS + warp * kDK + d], ctx[d]);
        }
      }
#pragma unroll
      for (int d = 0; d < kDK; ++d) Qs[lane * kLdS + warp * kDK + d] = ctx[d];   // only this lane ever read this Q row
    }
    group_sync(grp);
    // ---- (4) output projection + residual: warp = group of 4 channels, lane = row
    if (lane < L) {
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int c = warp * 4 + u; acc[u] = (c < D) ? EA[kAttnBO + c] + Xs[lane * kDP + c] : 0.f; }
      for (int k = 0; k < kHD; ++k) {
        const float cv = Qs[lane * kLdS + k];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(cv, EA[kAttnWO + (warp * 4 + u) * kHD + k], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) Ys[lane * kDP + warp * 4 + u] = acc[u];
    }
    group_sync(grp);
    // ---- LayerNorm, FFN, LayerNorm per row (warp 0, lane = row)
    if (warp == 0 && lane < L) {
      float y[kDP];
#pragma unroll
      for (int c = 0; c < kDP; ++c) y[c] = Ys[lane * kDP + c];
      layer_norm_inplace(y, D, EA + kAttnG, EA + kAttnB);
      float o[kDP];
#pragma unroll
      for (int c = 0; c < kDP; ++c) o[c] = (c < D) ? EF[kFfnB2 + c] + y[c] : 0.f;
      for (int mm = 0; mm < FF; ++mm) {
        float h = EF[kFfnB1 + mm];
#pragma unroll
        for (int c = 0; c < kDP; ++c) h = fmaf(y[c], EF[kFfnW1 + mm * kDP + c], h);
        h = fmaxf(h, 0.f);
#pragma unroll
        for (int c = 0; c < kDP; ++c) o[c] = fmaf(h, EF[kFfnW2 + c * kFFMax + mm], o[c]);
      }
      layer_norm_inplace(o, D, EF + kFfnG, EF + kFfnB);
#pragma unroll
      for (int c = 0; c < kDP; ++c) Es[lane * kDP + c] = o[c];
    }
    group_sync(grp);
    // ---- (5) decoder: K, V of the encoder output, Q of the query: thread = projection column
    {
      float wk[kDP], wv[kDP];
      float qd = DA[kAttnBQ + tid];
#pragma unroll
      for (int c = 0; c < kDP; ++c) {
        wk[c] = DA[kAttnWK + tid * kDP + c]; wv[c] = DA[kAttnWV + tid * kDP + c];
        qd = fmaf(qv[c], DA[kAttnWQ + tid * kDP + c], qd);
      }
      const float bk = DA[kAttnBK + tid], bv = DA[kAttnBV + tid];
      for (int i = 0; i < L; ++i) {
        float k = bk, v = bv;
#pragma unroll
        for (int c = 0; c < kDP; ++c) { const float e = Es[i * kDP + c]; k = fmaf(e, wk[c], k); v = fmaf(e, wv[c], v); }
        Ks[i * kLdS + tid] = k; Vs[i * kLdS + tid] = v;
      }
      Qs[tid] = qd;   // row 0 of Qs
    }
    group_sync(grp);
    {  // cross-attention: warp = head, lane = key; softmax across the warp; context: lane = channel of the head
      float a = -INFINITY;
      if (lane < L) {
        a = 0.f;
#pragma unroll
        for (int d = 0; d < kDK; ++d) a = fmaf(Qs[warp * kDK + d], Ks[lane * kLdS + warp * kDK + d], a);
        a *= p.scale;
      }
      const float m = warp_max(a);
      const float e = (lane < L) ? __expf(a - m) : 0.f;
      const float pj = e / warp_sum(e);
      float ctx = 0.f;
      for (int j = 0; j < L; ++j) ctx = fmaf(__shfl_sync(0xffffffffu, pj, j), Vs[j * kLdS + warp * kDK + lane], ctx);
      __syncwarp();
      Qs[kLdS + warp * kDK + lane] = ctx;   // row 1 of Qs
    }
    group_sync(grp);
    if (warp == 0) {  // output projection + residual + LayerNorm + FFN + LayerNorm for the single decoder row: lane = channel
      float o = 0.f;
      if (lane < D) {
        o = DA[kAttnBO + lane] + qv[lane];
        for (int k = 0; k < kHD; ++k) o = fmaf(Qs[kLdS + k], DA[kAttnWO + lane * kHD + k], o);
      }
      const float fD = static_cast<float>(D);
      float mean = warp_sum(lane < D ? o : 0.f) / fD;
      float dv = lane < D ? o - mean : 0.f;
      float rstd = rsqrtf(warp_sum(dv * dv) / fD + 1e-5f);
      const float dn = lane < D ? dv * rstd * DA[kAttnG + lane] + DA[kAttnB + lane] : 0.f;
      // FFN: lane = hidden unit (two per lane when d_ff > 32)
      float h0 = 0.f, h1 = 0.f;
      {
        float a0 = lane < FF ? DF[kFfnB1 + lane] : 0.f, a1 = lane + 32 < FF ? DF[kFfnB1 + lane + 32] : 0.f;
        for (int c = 0; c < D; ++c) {
          const float dc = __shfl_sync(0xffffffffu, dn, c);
          if (lane < FF) a0 = fmaf(dc, DF[kFfnW1 + lane * kDP + c], a0);
          if (lane + 32 < FF) a1 = fmaf(dc, DF[kFfnW1 + (lane + 32) * kDP + c], a1);
        }
        h0 = lane < FF ? fmaxf(a0, 0.f) : 0.f;
        h1 = lane + 32 < FF ? fmaxf(a1, 0.f) : 0.f;
      }
      float o2 = lane < D ? DF[kFfnB2 + lane] + dn : 0.f;
      for (int mm = 0; mm < 32; ++mm) {
        const float ha = __shfl_sync(0xffffffffu, h0, mm), hb = __shfl_sync(0xffffffffu, h1, mm);
        if (lane < D) {
          o2 = fmaf(ha, DF[kFfnW2 + lane * kFFMax + mm], o2);
          o2 = fmaf(hb, DF[kFfnW2 + lane * kFFMax + mm + 32], o2);
        }
      }
      mean = warp_sum(lane < D ? o2 : 0.f) / fD;
      dv = lane < D ? o2 - mean : 0.f;
      rstd = rsqrtf(warp_sum(dv * dv) / fD + 1e-5f);
      if (lane < D) out[t * D + lane] = dv * rstd * DF[kFfnG + lane] + DF[kFfnB + lane];
    }
    group_sync(grp);   // the tiles are rewritten by the next frame
  }
}

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

}  // namespace
}  // namespace sv

struct sv_trans {
  sv_trans_cfg cfg;
  int device = 0;
  std::map<std::string, sv::HostTensor> tensors;
  bool packed = false;
  float* d_blob = nullptr;
  int64_t* d_offsets = nullptr;
  int offsets_cap = 0;
};

namespace sv {
namespace {

// nn.Linear weight [rows, cols] (+ optional bias [rows]) -> blob[w_off + r * ld + c], blob[b_off + r]
int put_linear(const sv_trans* h, const std::string& p, int rows, int cols, int ld, std::vector<float>* blob, int w_off, int b_off) {
  auto it = h->tensors.find(p + ".weight");
  if (it == h->tensors.end()) return fail(SV_ERR_STATE, "trans: missing state_dict key '" + p + ".weight'");
  if (it->second.shape != std::vector<int64_t>{rows, cols}) return fail(SV_ERR_INVALID, "trans: wrong shape for '" + p + ".weight'");
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) (*blob)[w_off + r * ld + c] = it->second.data[static_cast<size_t>(r) * cols + c];
  auto ib = h->tensors.find(p + ".bias");
  if (ib != h->tensors.end()) {   // bias-free variants of the upstream code: treated as zero
    if (ib->second.shape != std::vector<int64_t>{rows}) return fail(SV_ERR_INVALID, "trans: wrong shape for '" + p + ".bias'");
    for (int r = 0; r < rows; ++r) (*blob)[b_off + r] = ib->second.data[r];
  }
  return SV_OK;
}
int put_norm(const sv_trans* h, const std::string& p, int D, std::vector<float>* blob, int g_off, int b_off) {
  auto ig = h->tensors.find(p + ".weight"), ib = h->tensors.find(p + ".bias");
  if (ig == h->tensors.end() || ib == h->tensors.end()) return fail(SV_ERR_STATE, "trans: missing LayerNorm '" + p + "'");
  if (ig->second.shape != std::vector<int64_t>{D} || ib->second.shape != std::vector<int64_t>{D}) return fail(SV_ERR_INVALID, "trans: wrong shape for '" + p + "'");
  for (int c = 0; c < D; ++c) { (*blob)[g_off + c] = ig->second.data[c]; (*blob)[b_off + c] = ib->second.data[c]; }
  return SV_OK;
}

}  // namespace
}  // namespace sv

extern "C" {

int sv_trans_create(const sv_trans_cfg* cfg, sv_trans_handle** out) {
  using namespace sv;
  SV_CHECK(cfg && out, "null argument");
  SV_CHECK(cfg->d_model >= 1 && cfg->d_model <= kDP, "trans: 1 <= d_model <= 16");
  SV_CHECK(cfg->len_q >= 1 && cfg->len_q <= kLMax, "trans: 1 <= len_q <= 32");
  SV_CHECK(cfg->d_ff >= 1 && cfg->d_ff <= kFFMax, "trans: 1 <= d_ff <= 64");
  if (cfg->n_heads != kH || cfg->d_k != kDK || cfg->d_v != kDK)
    return fail(SV_ERR_UNSUPPORTED, "trans: n_heads = 4 and d_k = d_v = 32 are supported (the reference's mstcn_f_maps = 32 configuration)");
  if (cfg->n_layers != 1) return fail(SV_ERR_UNSUPPORTED, "trans: n_layers must be 1 (adapter_transformer.py:322)");
  int dev = 0;
  SV_CUDA_OK(cudaGetDevice(&dev));
  sv_trans* h = new sv_trans();
  h->cfg = *cfg;
  h->device = dev;
  *out = h;
  return SV_OK;
}

int sv_trans_destroy(sv_trans_handle* h) {
  if (!h) return SV_OK;
  if (h->d_blob) cudaFree(h->d_blob);
  if (h->d_offsets) cudaFree(h->d_offsets);
  delete h;
  return SV_OK;
}

int sv_trans_set_tensor(sv_trans_handle* h, const char* name, const float* host_data, const int64_t* shape, int32_t ndim) {
  using namespace sv;
  SV_CHECK(h && name && host_data && (shape || ndim == 0), "null argument");
  HostTensor t;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= shape[i]; }
  t.data.assign(host_data, host_data + n);
  h->tensors[name] = std::move(t);
  h->packed = false;
  return SV_OK;
}

int sv_trans_pack_weights(sv_trans_handle* h) {
  using namespace sv;
  SV_CHECK(h, "null handle");
  const int D = h->cfg.d_model, FF = h->cfg.d_ff;
  std::vector<float> blob(kBlobSize, 0.f);
  const char* attn_name[2] = {"encoder.layers.0.enc_self_attn", "decoder.layers.0.dec_enc_attn"};
  const char* ffn_name[2] = {"encoder.layers.0.pos_ffn", "decoder.layers.0.pos_ffn"};
  for (int s = 0; s < 2; ++s) {
    const int a0 = s * (kAttnSize + kFfnSize), f0 = a0 + kAttnSize;
    const std::string ap = attn_name[s], fp = ffn_name[s];
    SV_TRY(put_linear(h, ap + ".W_Q", kHD, D, kDP, &blob, a0 + kAttnWQ, a0 + kAttnBQ));
    SV_TRY(put_linear(h, ap + ".W_K", kHD, D, kDP, &blob, a0 + kAttnWK, a0 + kAttnBK));
    SV_TRY(put_linear(h, ap + ".W_V", kHD, D, kDP, &blob, a0 + kAttnWV, a0 + kAttnBV));
    SV_TRY(put_linear(h, ap + ".fc", D, kHD, kHD, &blob, a0 + kAttnWO, a0 + kAttnBO));
    SV_TRY(put_norm(h, ap + ".layer_norm", D, &blob, a0 + kAttnG, a0 + kAttnB));
    SV_TRY(put_linear(h, fp + ".fc1", FF, D, kDP, &blob, f0 + kFfnW1, f0 + kFfnB1));
    SV_TRY(put_linear(h, fp + ".fc2", D, FF, kFFMax, &blob, f0 + kFfnW2, f0 + kFfnB2));
    SV_TRY(put_norm(h, fp + ".layer_norm", D, &blob, f0 + kFfnG, f0 + kFfnB));
  }
  SV_CUDA_OK(cudaSetDevice(h->device));
  if (!h->d_blob) SV_CUDA_OK(cudaMalloc(&h->d_blob, kBlobSize * sizeof(float)));
  SV_CUDA_OK(cudaMemcpy(h->d_blob, blob.data(), kBlobSize * sizeof(float), cudaMemcpyHostToDevice));
  h->packed = true;
  return SV_OK;
}

int sv_trans_forward(sv_trans_handle* h, const float* logits, int64_t ldx, const float* query, const int64_t* video_offsets, int32_t n_videos,
                     float* out, void* stream) {
  using namespace sv;
  SV_CHECK(h && logits && query && video_offsets && out, "null argument");
  if (!h->packed) return fail(SV_ERR_STATE, "trans: pack_weights() has not been called");
  SV_CHECK(n_videos >= 1 && video_offsets[0] == 0, "trans: offsets");
  for (int i = 0; i < n_videos; ++i) SV_CHECK(video_offsets[i + 1] > video_offsets[i], "trans: empty video / non-increasing offsets");
  const int64_t T = video_offsets[n_videos];
  SV_CHECK(ldx >= T, "trans: logits row stride");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->offsets_cap < n_videos + 1) {
    if (h->d_offsets) cudaFree(h->d_offsets);
    h->offsets_cap = std::max(128, 2 * (n_videos + 1));
    SV_CUDA_OK(cudaMalloc(&h->d_offsets, h->offsets_cap * sizeof(int64_t)));
  }
  SV_CUDA_OK(cudaMemcpyAsync(h->d_offsets, video_offsets, (n_videos + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  TransParams p;
  p.D = h->cfg.d_model; p.L = h->cfg.len_q; p.FF = h->cfg.d_ff; p.n_videos = n_videos;
  p.scale = 1.0f / sqrtf(static_cast<float>(h->cfg.d_k));
  const int smem = (kBlobSize + kGroups * kGroupFloats) * static_cast<int>(sizeof(float));
  SV_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(trans_head_kernel), smem));
  const int grid = static_cast<int>(std::min<int64_t>((T + kGroups - 1) / kGroups, std::max(1, device_sm_count())));
  trans_head_kernel<<<grid, 128 * kGroups, smem, st>>>(logits, ldx, query, h->d_offsets, T, h->d_blob, out, p);
  return launch_status("trans_head_kernel");
}

}  // extern "C"
