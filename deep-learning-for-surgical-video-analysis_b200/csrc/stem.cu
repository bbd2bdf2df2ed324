// First-layer convolutions of the three input branches, fused end to end:
//   patch_embed1          Conv2d(3 -> 64, k7 s4 p3) + LayerNorm(1e-5)          (mix_transformer_evp.py:228-229, 209-215)
//   handcrafted_generator1 Conv2d(3 -> 16, k7 s4 p3) + LayerNorm(1e-5)          (:582-583, on the Gaussian-filtered segmap)
//   flow_encoder.conv1    Conv2d(2 -> 64, k7 s4 p3) + BatchNorm(eval) + ReLU   (:823-825, :846; BN folded into the weights)
// The inputs are fp32 NCHW with 2-3 channels, so K = 49*Cin is tiny (98 / 147) and a materialised im2col buffer would cost
// more HBM traffic than the image itself.  One CTA = one 8x8 tile of output pixels: the 35x35xCin input patch is staged in
// shared memory (bf16), expanded there into the [64 x K] operand, multiplied on tensor cores (mma.sync m16n8k16 — the
// contraction is 0.4 % of the path's FLOPs) and normalised / activated in registers before the only global writes.
#include <mutex>

#include "kernels.cuh"

namespace sv {
namespace {

constexpr int kTileP = 8;                   // 8 x 8 output pixels per tile
constexpr int kPatch = kTileP * 4 + 3;      // 35 x 35 input pixels (k7, stride 4)
constexpr int kKPad = 176;                  // in-kernel K order is (c, kh, kw padded 7 -> 8): Cin*56 (<= 168) rounded up to 16
constexpr int kLdA = kKPad + 8;             // smem row stride: odd multiple of 16 B -> conflict-free ldmatrix
constexpr int kMaxLdw = 160;                // packed global weights: [Cout][ldw], k = (kh, kw, cin), ldw <= 160

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct StemParams {
  const float* src;      // [B, Cin, H, W] fp32
  const bf16* w;         // [COUT, ldw] bf16, k = (kh, kw, cin), zero padded
  const float* bias;     // [COUT]
  const float* gamma;    // LayerNorm affine (mode 0) or unused
  const float* beta;
  float* out_f32;        // [B*Ho*Wo, COUT] or null
  bf16* out_bf16;        // [B*Ho*Wo, COUT] or null
  int B, Cin, H, W, Ho, Wo, ldw;
  int tiles_x, tiles_y;
  int mode;              // 0: + bias, LayerNorm(eps); 1: + bias, ReLU
  float eps;
};

template <int COUT>
__global__ void __launch_bounds__(128) stem_conv_kernel(const StemParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Ws = reinterpret_cast<bf16*>(smem_raw);                 // [COUT][kLdA]
  bf16* As = Ws + COUT * kLdA;                                  // [64][kLdA]
  bf16* Ps = As + 64 * kLdA;                                    // [Cin][35][40] input patch (row stride 80 B keeps 8-byte alignment)
  constexpr int kPs = kPatch + 5;                               // patch row stride (elements)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ksteps = (p.Cin * 56 + 15) / 16;                    // k-steps actually needed (7 for Cin=2, 11 for Cin=3)

  // weights -> smem, re-ordered from the packed (kh, kw, c) layout to (c, kh, kw8); the kw = 7 slot and the K tail are zero,
  // so whatever the operand holds there is multiplied by 0
  for (int i = tid; i < COUT * kKPad; i += 128) {
    const int n = i / kKPad, kk = i % kKPad;
    const int kw = kk & 7, ch = kk >> 3;                        // ch = c*7 + kh
    const int c = ch / 7, kh = ch % 7;
    bf16 v = __float2bfloat16(0.f);
    if (kw < 7 && c < p.Cin) v = p.w[static_cast<int64_t>(n) * p.ldw + (kh * 7 + kw) * p.Cin + c];
    Ws[n * kLdA + kk] = v;
  }
  for (int i = tid; i < 64 * (kLdA - p.Cin * 56) ; i += 128) {  // zero the K tail of the operand tile once (never rewritten)
    const int w = kLdA - p.Cin * 56;
    As[(i / w) * kLdA + p.Cin * 56 + (i % w)] = __float2bfloat16(0.f);
  }
  __syncthreads();

  const int tiles_per_frame = p.tiles_x * p.tiles_y;
  const long long num_tiles = static_cast<long long>(p.B) * tiles_per_frame;
  for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int b = static_cast<int>(t / tiles_per_frame);
    const int tr = static_cast<int>(t % tiles_per_frame);
    const int oy0 = (tr / p.tiles_x) * kTileP, ox0 = (tr % p.tiles_x) * kTileP;
    const int iy0 = oy0 * 4 - 3, ix0 = ox0 * 4 - 3;
    // ---- input patch -> smem (bf16), zero outside the image (= conv zero padding); warp = patch rows, lane = columns
    const float* sb = p.src + static_cast<int64_t>(b) * p.Cin * p.H * p.W;
    // warp = patch rows (yy = warp, warp + 4, ...), lane = patch columns: 32 consecutive floats of an image row per load (+ a 3-pixel
    // tail), no per-element div/mod; the 9 row loads of a channel are issued before any is consumed (~9 x 128 B in flight per warp)
    for (int c = 0; c < p.Cin; ++c) {
      const float* sc = sb + static_cast<int64_t>(c) * p.H * p.W;
      bf16* pc = Ps + c * kPatch * kPs;
      float v0[9], v1[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int yy = warp + 4 * j, iy = iy0 + yy;
        const bool rok = yy < kPatch && iy >= 0 && iy < p.H;
        const int ix = ix0 + lane, ix2 = ix0 + 32 + lane;
        v0[j] = (rok && ix >= 0 && ix < p.W) ? __ldg(sc + static_cast<int64_t>(iy) * p.W + ix) : 0.f;
        v1[j] = (rok && lane < kPatch - 32 && ix2 < p.W) ? __ldg(sc + static_cast<int64_t>(iy) * p.W + ix2) : 0.f;   // ix2 >= 29 >= 0
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int yy = warp + 4 * j;
        if (yy < kPatch) {
          pc[yy * kPs + lane] = __float2bfloat16(v0[j]);
          if (lane < kPs - 32) pc[yy * kPs + 32 + lane] = __float2bfloat16(v1[j]);   // columns 35..39 of the padded row are zero
        }
      }
    }
    __syncthreads();
    // ---- expand to the [64 x K] operand: row m = (py, px); for every (c, kh) copy 8 consecutive patch pixels (kw 0..7)
    const int nch = p.Cin * 7;
    for (int i = tid; i < 64 * nch; i += 128) {
      const int m = i & 63, ch = i >> 6;                        // consecutive threads -> consecutive pixels
      const int c = ch / 7, kh = ch - c * 7;
      const bf16* src = Ps + (c * kPatch + (m >> 3) * 4 + kh) * kPs + (m & 7) * 4;   // 8-byte aligned
      const uint2 lo = *reinterpret_cast<const uint2*>(src);
      const uint2 hi = *reinterpret_cast<const uint2*>(src + 4);
      *reinterpret_cast<uint4*>(As + m * kLdA + ch * 8) = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
    __syncthreads();
    // ---- tensor-core contraction: warp = 16 pixels (two rows of the tile) x COUT
    float acc[COUT / 8][4];
#pragma unroll
    for (int i = 0; i < COUT / 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const int mi = lane >> 3;
    for (int ks = 0; ks < ksteps; ++ks) {
      uint32_t af[4];
      ldsm_x4(af, static_cast<uint32_t>(__cvta_generic_to_shared(As + (warp * 16 + (mi & 1) * 8 + (lane & 7)) * kLdA + ks * 16 + (mi >> 1) * 8)));
#pragma unroll
      for (int np = 0; np < COUT / 16; ++np) {
        uint32_t bfr[4];
        ldsm_x4(bfr, static_cast<uint32_t>(__cvta_generic_to_shared(Ws + (np * 16 + (mi >> 1) * 8 + (lane & 7)) * kLdA + ks * 16 + (mi & 1) * 8)));
        mma16816(acc[np * 2], af, bfr[0], bfr[1]);
        mma16816(acc[np * 2 + 1], af, bfr[2], bfr[3]);
      }
    }
    // ---- epilogue in registers: thread holds rows g, g+8 of the warp's 16 pixels, columns nt*8 + 2t, +1
    const int g = lane >> 2, tq = lane & 3;
    float mean[2] = {0.f, 0.f}, rstd[2] = {1.f, 1.f};
#pragma unroll
    for (int nt = 0; nt < COUT / 8; ++nt) {
      const float2 bb = __ldg(reinterpret_cast<const float2*>(p.bias + nt * 8 + tq * 2));
      acc[nt][0] += bb.x; acc[nt][1] += bb.y; acc[nt][2] += bb.x; acc[nt][3] += bb.y;
    }
    if (p.mode == 0) {
      float s[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < COUT / 8; ++nt) { s[0] += acc[nt][0] + acc[nt][1]; s[1] += acc[nt][2] + acc[nt][3]; }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        s[r] += __shfl_xor_sync(0xffffffffu, s[r], 1);
        s[r] += __shfl_xor_sync(0xffffffffu, s[r], 2);
        mean[r] = s[r] * (1.0f / COUT);
      }
      float q[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < COUT / 8; ++nt) {
        const float a = acc[nt][0] - mean[0], bq = acc[nt][1] - mean[0], c = acc[nt][2] - mean[1], d = acc[nt][3] - mean[1];
        q[0] += a * a + bq * bq;
        q[1] += c * c + d * d;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        q[r] += __shfl_xor_sync(0xffffffffu, q[r], 1);
        q[r] += __shfl_xor_sync(0xffffffffu, q[r], 2);
        rstd[r] = 1.0f / sqrtf(q[r] * (1.0f / COUT) + p.eps);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int m = warp * 16 + g + r * 8;
      const int oy = oy0 + (m >> 3), ox = ox0 + (m & 7);
      if (oy < p.Ho && ox < p.Wo) {
        const int64_t row = (static_cast<int64_t>(b) * p.Ho + oy) * p.Wo + ox;
#pragma unroll
        for (int nt = 0; nt < COUT / 8; ++nt) {
          const int col = nt * 8 + tq * 2;
          float v0 = acc[nt][r * 2], v1 = acc[nt][r * 2 + 1];
          if (p.mode == 0) {
            const float2 gm = __ldg(reinterpret_cast<const float2*>(p.gamma + col));
            const float2 bt = __ldg(reinterpret_cast<const float2*>(p.beta + col));
            v0 = (v0 - mean[r]) * rstd[r] * gm.x + bt.x;
            v1 = (v1 - mean[r]) * rstd[r] * gm.y + bt.y;
          } else {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          }
          if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + row * COUT + col) = make_float2(v0, v1);
          if (p.out_bf16) *reinterpret_cast<uint32_t*>(p.out_bf16 + row * COUT + col) = pack_bf16x2(v0, v1);
        }
      }
    }
    __syncthreads();  // As / Ps are rewritten by the next tile
  }
}

template <int COUT>
int stem_launch(const StemParams& p, cudaStream_t st) {
  constexpr size_t smem = (static_cast<size_t>(COUT) * kLdA + 64 * kLdA + 3 * kPatch * (kPatch + 5)) * sizeof(bf16);
  SV_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(stem_conv_kernel<COUT>), static_cast<int>(smem)));
  const long long tiles = static_cast<long long>(p.B) * p.tiles_x * p.tiles_y;
  const int grid = static_cast<int>(std::min<long long>(tiles, 8LL * device_sm_count()));
  stem_conv_kernel<COUT><<<grid, 128, smem, st>>>(p);
  return launch_status("stem_conv_kernel");
}

}  // namespace

int launch_stem_conv(const float* src, const bf16* w, int ldw, const float* bias, const float* gamma, const float* beta, float eps, int mode,
                     int B, int Cin, int H, int W, int Cout, float* out_f32, bf16* out_bf16, cudaStream_t st) {
  SV_CHECK(Cin >= 1 && Cin <= 3, "stem conv supports 1..3 input channels");
  SV_CHECK(ldw % 8 == 0 && ldw >= 49 * Cin && ldw <= kMaxLdw, "stem conv weight row stride");
  SV_CHECK(mode == 1 || (gamma && beta), "stem conv LayerNorm mode needs gamma/beta");
  StemParams p;
  p.src = src; p.w = w; p.bias = bias; p.gamma = gamma; p.beta = beta; p.out_f32 = out_f32; p.out_bf16 = out_bf16;
  p.B = B; p.Cin = Cin; p.H = H; p.W = W; p.Ho = conv_out_dim(H, 7, 4, 3); p.Wo = conv_out_dim(W, 7, 4, 3); p.ldw = ldw;
  p.tiles_x = ceil_div(p.Wo, kTileP); p.tiles_y = ceil_div(p.Ho, kTileP); p.mode = mode; p.eps = eps;
  SV_CHECK(p.Ho > 0 && p.Wo > 0, "stem conv output empty");
  if (Cout == 64) return stem_launch<64>(p, st);
  if (Cout == 16) return stem_launch<16>(p, st);
  if (Cout == 32) return stem_launch<32>(p, st);
  return fail(SV_ERR_UNSUPPORTED, "stem conv supports 16/32/64 output channels");
}

bool stem_conv_supported(int Cin, int Cout, int ldw) { return Cin <= 3 && (Cout == 64 || Cout == 32 || Cout == 16) && ldw <= kMaxLdw; }

}  // namespace sv

extern "C" int sv_op_stem_conv(const float* src, const uint16_t* w, int32_t ldw, const float* bias, const float* gamma, const float* beta,
                               float eps, int32_t relu, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout, float* out_f32,
                               uint16_t* out_bf16, void* stream) {
  return sv::launch_stem_conv(src, reinterpret_cast<const sv::bf16*>(w), ldw, bias, gamma, beta, eps, relu ? 1 : 0, B, Cin, H, W, Cout, out_f32,
                              reinterpret_cast<sv::bf16*>(out_bf16), static_cast<cudaStream_t>(stream));
}
