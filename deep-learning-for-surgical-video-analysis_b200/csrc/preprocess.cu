// On-GPU input transforms: the per-frame CPU work of the reference's dataset class done on the device,
// so that uint8 frames (and the raw RAFT flow) cross PCIe instead of 1.6 MB/frame of normalised fp32.
//
//   image / segmentation map, uint8 [B, H, W, 3]:
//       transforms.Resize((R, R)) on a PIL image  ->  CenterCrop(C)  ->  ToTensor()  ->  Normalize(mean, std)
//       (generate_evp_LFB.py:242-248).  Pillow's resize is a separable antialiased bilinear filter with 22-bit fixed-point
//       coefficients and a uint8 intermediate between the horizontal and the vertical pass; both passes are reproduced
//       with the same integer arithmetic, so the result is bit-identical to torchvision's (tests/test_preprocess_gpu.py).
//   flow, float32 [B, H, W, 2]:
//       cv2.resize(flow, (R, R), INTER_LINEAR); u *= R/W; v *= R/H; -> [2, R, R] -> CenterCrop(C)   (data_process.py:432-481)
//       float32 lerps with every product rounded (no FMA contraction), matching OpenCV's scalar arithmetic.
//
// Only the source rows/columns that the centre crop can see are touched.  The coefficient tables are built once per
// geometry on the host, exactly as the two libraries build them, and live in the handle.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace sv {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow Resample.c

struct AxisU8 {       // antialiased bilinear taps for the cropped output range of one axis
  std::vector<int> first;   // first source index per output
  std::vector<int> count;   // taps per output
  std::vector<int> kk;      // [n_out, ksize] fixed point
  int ksize = 0;
  int src_lo = 0, src_hi = 0;  // source range [lo, hi) the outputs touch
};

// Resample.c precompute_coeffs (bilinear, support 1) + normalize_coeffs_8bpc for outputs [o0, o0 + n) of an axis in -> out
AxisU8 pil_axis(int in_size, int out_size, int o0, int n) {
  AxisU8 a;
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  a.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  a.first.resize(n);
  a.count.resize(n);
  a.kk.assign(static_cast<size_t>(n) * a.ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(a.ksize);
  a.src_lo = in_size;
  a.src_hi = 0;
  for (int i = 0; i < n; ++i) {
    const int xx = o0 + i;
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0) t = -t;
      w[x] = t < 1.0 ? 1.0 - t : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) w[x] /= ww;
      a.kk[static_cast<size_t>(i) * a.ksize + x] =
          w[x] < 0 ? static_cast<int>(-0.5 + w[x] * (1 << kPrecisionBits)) : static_cast<int>(0.5 + w[x] * (1 << kPrecisionBits));
    }
    a.first[i] = xmin;
    a.count[i] = xmax;
    a.src_lo = std::min(a.src_lo, xmin);
    a.src_hi = std::max(a.src_hi, xmin + xmax);
  }
  return a;
}

struct AxisF32 {  // OpenCV INTER_LINEAR source pair + fraction for the cropped output range of one axis
  std::vector<int> i0, i1;
  std::vector<float> frac;
};

// resize.cpp: fx = (float)((d + 0.5) * scale - 0.5); s = cvFloor(fx); fx -= s; clamp to the borders
AxisF32 cv_axis(int in_size, int out_size, int o0, int n) {
  AxisF32 a;
  a.i0.resize(n); a.i1.resize(n); a.frac.resize(n);
  const double inv_scale = static_cast<double>(out_size) / in_size;
  const double scale = 1.0 / inv_scale;
  for (int i = 0; i < n; ++i) {
    const int d = o0 + i;
    float fx = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(std::floor(fx));
    fx -= static_cast<float>(s);
    if (in_size == out_size) { s = d; fx = 0.f; }  // cv2.resize returns a copy when the size is unchanged
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= in_size - 1) { s = in_size - 1; fx = 0.f; }
    a.i0[i] = s;
    a.i1[i] = std::min(s + 1, in_size - 1);
    a.frac[i] = fx;
  }
  return a;
}

__device__ __forceinline__ int clip8(int v) { return min(max(v >> kPrecisionBits, 0), 255); }

// horizontal pass: src u8 [B, H, W, 3] rows [row_lo, row_lo + rows) -> tmp u8 [B, rows, n_out, 3]
__global__ void prep_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp, const int* __restrict__ first,
                              const int* __restrict__ count, const int* __restrict__ kk, int ksize, int B, int H, int W, int row_lo,
                              int rows, int n_out) {
  const long long total = static_cast<long long>(B) * rows * n_out;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % n_out);
    const long long t = i / n_out;
    const int r = static_cast<int>(t % rows);
    const int b = static_cast<int>(t / rows);
    const int x0 = first[xx], n = count[xx];
    const int* k = kk + static_cast<size_t>(xx) * ksize;
    const uint8_t* s = src + ((static_cast<long long>(b) * H + row_lo + r) * W + x0) * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < n; ++x) {
      const int kv = k[x];
      a0 += s[x * 3 + 0] * kv;
      a1 += s[x * 3 + 1] * kv;
      a2 += s[x * 3 + 2] * kv;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = static_cast<uint8_t>(clip8(a0));
    o[1] = static_cast<uint8_t>(clip8(a1));
    o[2] = static_cast<uint8_t>(clip8(a2));
  }
}

// vertical pass + ToTensor + Normalize: tmp u8 [B, rows, n_out_x, 3] -> out f32 [B, 3, n_out_y, n_out_x]
__global__ void prep_v_kernel(const uint8_t* __restrict__ tmp, float* __restrict__ out, const int* __restrict__ first,
                              const int* __restrict__ count, const int* __restrict__ kk, int ksize, int B, int row_lo, int rows,
                              int n_out_y, int n_out_x, float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long total = static_cast<long long>(B) * n_out_y * n_out_x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % n_out_x);
    const long long t = i / n_out_x;
    const int yy = static_cast<int>(t % n_out_y);
    const int b = static_cast<int>(t / n_out_y);
    const int y0 = first[yy] - row_lo, n = count[yy];
    const int* k = kk + static_cast<size_t>(yy) * ksize;
    const uint8_t* s = tmp + ((static_cast<long long>(b) * rows + y0) * n_out_x + xx) * 3;
    const long long rstride = static_cast<long long>(n_out_x) * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < n; ++y) {
      const int kv = k[y];
      a0 += s[y * rstride + 0] * kv;
      a1 += s[y * rstride + 1] * kv;
      a2 += s[y * rstride + 2] * kv;
    }
    // ToTensor: float(u8) / 255; Normalize: (x - mean) / std — IEEE float32, one rounding per operation
    const float v0 = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(clip8(a0)), 255.0f), m0), s0);
    const float v1 = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(clip8(a1)), 255.0f), m1), s1);
    const float v2 = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(clip8(a2)), 255.0f), m2), s2);
    const long long plane = static_cast<long long>(n_out_y) * n_out_x;
    float* o = out + static_cast<long long>(b) * 3 * plane + static_cast<long long>(yy) * n_out_x + xx;
    o[0] = v0;
    o[plane] = v1;
    o[2 * plane] = v2;
  }
}

// flow: f32 [B, H, W, 2] -> f32 [B, 2, n_out_y, n_out_x]; horizontal lerp of the two source rows, vertical lerp, displacement rescale
__global__ void prep_flow_kernel(const float* __restrict__ src, float* __restrict__ out, const int* __restrict__ xi0, const int* __restrict__ xi1,
                                 const float* __restrict__ xf, const int* __restrict__ yi0, const int* __restrict__ yi1,
                                 const float* __restrict__ yf, int B, int H, int W, int n_out_y, int n_out_x, float scale_u, float scale_v) {
  const long long total = static_cast<long long>(B) * n_out_y * n_out_x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % n_out_x);
    const long long t = i / n_out_x;
    const int yy = static_cast<int>(t % n_out_y);
    const int b = static_cast<int>(t / n_out_y);
    const float a1 = xf[xx], a0 = __fsub_rn(1.0f, a1);
    const float b1 = yf[yy], b0 = __fsub_rn(1.0f, b1);
    const float2* img = reinterpret_cast<const float2*>(src) + static_cast<long long>(b) * H * W;
    const float2 p00 = __ldg(img + static_cast<long long>(yi0[yy]) * W + xi0[xx]);
    const float2 p01 = __ldg(img + static_cast<long long>(yi0[yy]) * W + xi1[xx]);
    const float2 p10 = __ldg(img + static_cast<long long>(yi1[yy]) * W + xi0[xx]);
    const float2 p11 = __ldg(img + static_cast<long long>(yi1[yy]) * W + xi1[xx]);
    const float r0u = __fadd_rn(__fmul_rn(p00.x, a0), __fmul_rn(p01.x, a1));
    const float r1u = __fadd_rn(__fmul_rn(p10.x, a0), __fmul_rn(p11.x, a1));
    const float r0v = __fadd_rn(__fmul_rn(p00.y, a0), __fmul_rn(p01.y, a1));
    const float r1v = __fadd_rn(__fmul_rn(p10.y, a0), __fmul_rn(p11.y, a1));
    const float u = __fmul_rn(__fadd_rn(__fmul_rn(r0u, b0), __fmul_rn(r1u, b1)), scale_u);
    const float v = __fmul_rn(__fadd_rn(__fmul_rn(r0v, b0), __fmul_rn(r1v, b1)), scale_v);
    const long long plane = static_cast<long long>(n_out_y) * n_out_x;
    float* o = out + static_cast<long long>(b) * 2 * plane + static_cast<long long>(yy) * n_out_x + xx;
    o[0] = u;
    o[plane] = v;
  }
}

template <typename T>
int upload(const std::vector<T>& v, T** dptr) {
  SV_CUDA_OK(cudaMalloc(dptr, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) SV_CUDA_OK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return SV_OK;
}

inline int grid_for(long long total, int threads) {
  const long long blocks = (total + threads - 1) / threads;
  return static_cast<int>(std::min<long long>(blocks, static_cast<long long>(device_sm_count()) * 16));
}

}  // namespace
}  // namespace sv

struct sv_prep {
  int in_h = 0, in_w = 0, flow_h = 0, flow_w = 0, resize = 0, crop = 0, top = 0, left = 0;
  float mean[3], stdv[3];
  // image tables
  int *x_first = nullptr, *x_count = nullptr, *x_kk = nullptr, x_ksize = 0;
  int *y_first = nullptr, *y_count = nullptr, *y_kk = nullptr, y_ksize = 0;
  int row_lo = 0, rows = 0;
  // flow tables
  int *fx0 = nullptr, *fx1 = nullptr, *fy0 = nullptr, *fy1 = nullptr;
  float *fxf = nullptr, *fyf = nullptr;
  float scale_u = 1.f, scale_v = 1.f;
};

extern "C" {

int sv_prep_create(int32_t in_h, int32_t in_w, int32_t flow_h, int32_t flow_w, int32_t resize, int32_t crop, const float* mean3,
                   const float* std3, sv_prep** out) {
  using namespace sv;
  SV_CHECK(out != nullptr && mean3 != nullptr && std3 != nullptr, "sv_prep_create: null argument");
  SV_CHECK(in_h > 0 && in_w > 0 && resize > 0 && crop > 0 && crop <= resize, "sv_prep_create: bad geometry");
  SV_CHECK((flow_h > 0 && flow_w > 0) || (flow_h == 0 && flow_w == 0), "sv_prep_create: flow size must be both set or both 0");
  sv_prep* h = new sv_prep();
  h->in_h = in_h; h->in_w = in_w; h->flow_h = flow_h; h->flow_w = flow_w; h->resize = resize; h->crop = crop;
  // torchvision center_crop: int(round((size - crop) / 2.0)) — Python round() is half-to-even
  h->top = h->left = static_cast<int>(std::nearbyint((resize - crop) / 2.0));
  for (int i = 0; i < 3; ++i) { h->mean[i] = mean3[i]; h->stdv[i] = std3[i]; }
  const AxisU8 ax = pil_axis(in_w, resize, h->left, crop);
  const AxisU8 ay = pil_axis(in_h, resize, h->top, crop);
  h->x_ksize = ax.ksize; h->y_ksize = ay.ksize; h->row_lo = ay.src_lo; h->rows = ay.src_hi - ay.src_lo;
  int rc = upload(ax.first, &h->x_first);
  if (rc == SV_OK) rc = upload(ax.count, &h->x_count);
  if (rc == SV_OK) rc = upload(ax.kk, &h->x_kk);
  if (rc == SV_OK) rc = upload(ay.first, &h->y_first);
  if (rc == SV_OK) rc = upload(ay.count, &h->y_count);
  if (rc == SV_OK) rc = upload(ay.kk, &h->y_kk);
  if (rc == SV_OK && flow_h > 0) {
    const AxisF32 fx = cv_axis(flow_w, resize, h->left, crop);
    const AxisF32 fy = cv_axis(flow_h, resize, h->top, crop);
    rc = upload(fx.i0, &h->fx0);
    if (rc == SV_OK) rc = upload(fx.i1, &h->fx1);
    if (rc == SV_OK) rc = upload(fx.frac, &h->fxf);
    if (rc == SV_OK) rc = upload(fy.i0, &h->fy0);
    if (rc == SV_OK) rc = upload(fy.i1, &h->fy1);
    if (rc == SV_OK) rc = upload(fy.frac, &h->fyf);
    // flow_resized[:, :, 0] *= target_w / w_origin (a Python float applied to a float32 array: float32 multiply)
    h->scale_u = static_cast<float>(static_cast<double>(resize) / flow_w);
    h->scale_v = static_cast<float>(static_cast<double>(resize) / flow_h);
  }
  if (rc != SV_OK) { sv_prep_destroy(h); return rc; }
  *out = h;
  return SV_OK;
}

int sv_prep_destroy(sv_prep* h) {
  if (!h) return SV_OK;
  void* ptrs[] = {h->x_first, h->x_count, h->x_kk, h->y_first, h->y_count, h->y_kk, h->fx0, h->fx1, h->fy0, h->fy1, h->fxf, h->fyf};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
  return SV_OK;
}

size_t sv_prep_workspace_bytes(const sv_prep* h, int32_t B) {
  if (!h || B <= 0) return 0;
  return static_cast<size_t>(B) * h->rows * h->crop * 3;
}

int sv_prep_images(sv_prep* h, const uint8_t* src, int32_t B, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace sv;
  SV_CHECK(h != nullptr && src != nullptr && out != nullptr, "sv_prep_images: null argument");
  if (B <= 0) return SV_OK;
  SV_CHECK(workspace != nullptr && workspace_bytes >= sv_prep_workspace_bytes(h, B), "sv_prep_images: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* tmp = static_cast<uint8_t*>(workspace);
  const long long n1 = static_cast<long long>(B) * h->rows * h->crop;
  prep_h_kernel<<<grid_for(n1, 256), 256, 0, st>>>(src, tmp, h->x_first, h->x_count, h->x_kk, h->x_ksize, B, h->in_h, h->in_w, h->row_lo, h->rows,
                                                   h->crop);
  SV_TRY(launch_status("prep_h_kernel"));
  const long long n2 = static_cast<long long>(B) * h->crop * h->crop;
  prep_v_kernel<<<grid_for(n2, 256), 256, 0, st>>>(tmp, out, h->y_first, h->y_count, h->y_kk, h->y_ksize, B, h->row_lo, h->rows, h->crop, h->crop,
                                                   h->mean[0], h->mean[1], h->mean[2], h->stdv[0], h->stdv[1], h->stdv[2]);
  return launch_status("prep_v_kernel");
}

int sv_prep_flow(sv_prep* h, const float* flow, int32_t B, float* out, void* stream) {
  using namespace sv;
  SV_CHECK(h != nullptr && flow != nullptr && out != nullptr, "sv_prep_flow: null argument");
  SV_CHECK(h->flow_h > 0, "sv_prep_flow: handle was created without a flow geometry");
  if (B <= 0) return SV_OK;
  const long long n = static_cast<long long>(B) * h->crop * h->crop;
  prep_flow_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(flow, out, h->fx0, h->fx1, h->fxf, h->fy0, h->fy1, h->fyf, B,
                                                                                     h->flow_h, h->flow_w, h->crop, h->crop, h->scale_u, h->scale_v);
  return launch_status("prep_flow_kernel");
}

}  // extern "C"
