// GPU input pipeline, image-shaped modalities.
//
// Replaces the per-sample torchvision chain of the reference's processors
//   image / thermal: ToTensor -> Resize(224, BICUBIC) -> CenterCrop(224) -> Normalize(OPENAI mean / std)
//                    (languagebind/image/processing_image.py:20-29, thermal/processing_thermal.py:15-25)
//   depth:           DepthNorm (/1000, clip [0.01, max_depth], / max_depth, 1 -> 3 channels) -> the same three
//                    (languagebind/depth/processing_depth.py:21-57)
// which the unchanged loader runs on the host for every sample (src/dataset/data_loader.py:74-78, 0 workers).
// ONE kernel per decoded image: reads the uint8 HWC (or float HW) pixels once, writes the normalised fp32 CHW crop
// once -- nothing of the resized intermediate is materialised, and only the S x S crop is ever computed.
//
// Resampling follows torch's F.interpolate(mode='bicubic', align_corners=False), which torchvision's Resize calls on
// tensors: antialias = 0 is the 4-tap cubic convolution (A = -0.75, border indices clamped) -- torchvision <= 0.16
// default on tensors, the reference's era; antialias = 1 is the area-scaled separable filter (A = -0.5, support
// 2 * scale, window truncated at the border, weights renormalised) -- torchvision >= 0.17 default.
// Each 16 x 16 output tile first builds its 16 column and 16 row tap tables (index + weight) in shared memory.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

constexpr int kTile = 16;

__device__ __forceinline__ float cubic_aa(float x) {               // A = -0.5 (aten: bicubic anti-alias filter)
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}
__device__ __forceinline__ float cubic_c1(float x, float A) { return ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f; }
__device__ __forceinline__ float cubic_c2(float x, float A) { return ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A; }

// taps of output coordinate `o` (in the RESIZED image) along an axis of in_size -> out_size; returns the tap count
__device__ int fill_taps(int in_size, int out_size, int o, int antialias, int max_taps, int* idx, float* w) {
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  if (!antialias) {
    // ONE rounding (fused multiply-add), as aten's own kernels compute it: at coordinates ~10^3 a separately rounded
    // product moves the fractional offset by 6e-5 -- visible at the 1e-5 parity bar
    const float src = fmaf(scale, static_cast<float>(o) + 0.5f, -0.5f);
    const float fl = floorf(src);
    const float t = src - fl;
    const int f = static_cast<int>(fl);
    const float A = -0.75f;
    const float c[4] = {cubic_c2(t + 1.0f, A), cubic_c1(t, A), cubic_c1(1.0f - t, A), cubic_c2(2.0f - t, A)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = f - 1 + j;
      k = k < 0 ? 0 : (k > in_size - 1 ? in_size - 1 : k);
      idx[j] = k, w[j] = c[j];
    }
    return 4;
  }
  const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const float center = scale * (static_cast<float>(o) + 0.5f);
  int xmin = static_cast<int>(center - support + 0.5f);
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(center + support + 0.5f);
  if (xmax > in_size) xmax = in_size;
  int n = xmax - xmin;
  if (n > max_taps) n = max_taps;                                    // (the host refuses scales that would get here)
  float total = 0.0f;
  for (int j = 0; j < n; ++j) {
    const float v = cubic_aa((static_cast<float>(j + xmin) - center + 0.5f) * invscale);
    idx[j] = xmin + j, w[j] = v;
    total += v;
  }
  const float inv = total != 0.0f ? 1.0f / total : 0.0f;
  for (int j = 0; j < n; ++j) w[j] *= inv;
  return n;
}

struct PreprocParams {
  const void* src;
  float* dst;
  int src_f32, H, W, RH, RW, top, left, S, antialias, max_taps;
  float pre_div, lo, hi, post_div;
  float mean[3], std_[3];
};

template <bool F32>
__global__ void __launch_bounds__(kTile* kTile) image_preprocess_kernel(const PreprocParams p) {
  extern __shared__ unsigned char smem_raw[];
  const int T = p.max_taps;
  int* xi = reinterpret_cast<int*>(smem_raw);                        // [kTile][T]
  int* yi = xi + kTile * T;
  float* xw = reinterpret_cast<float*>(yi + kTile * T);
  float* yw = xw + kTile * T;
  __shared__ int xn[kTile], yn[kTile];
  __shared__ float lut[256];                                         // ToTensor: byte / 255, the division done once
  if (!F32) lut[threadIdx.x] = static_cast<float>(threadIdx.x) / p.pre_div;   // (per value, not per tap: the IEEE
  const int tx = threadIdx.x % kTile, ty = threadIdx.x / kTile;      //  divides were 3/4 of the antialiased kernel)
  const int ox0 = blockIdx.x * kTile, oy0 = blockIdx.y * kTile;
  if (threadIdx.x < kTile) {
    const int ox = ox0 + threadIdx.x;
    xn[threadIdx.x] = ox < p.S ? fill_taps(p.W, p.RW, ox + p.left, p.antialias, T, xi + threadIdx.x * T, xw + threadIdx.x * T) : 0;
  } else if (threadIdx.x < 2 * kTile) {
    const int r = threadIdx.x - kTile;
    const int oy = oy0 + r;
    yn[r] = oy < p.S ? fill_taps(p.H, p.RH, oy + p.top, p.antialias, T, yi + r * T, yw + r * T) : 0;
  }
  __syncthreads();
  const int ox = ox0 + tx, oy = oy0 + ty;
  if (ox >= p.S || oy >= p.S) return;
  const int nx = xn[tx], ny = yn[ty];
  const int* xidx = xi + tx * T;
  const float* xwt = xw + tx * T;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  for (int i = 0; i < ny; ++i) {
    const long row = static_cast<long>(yi[ty * T + i]) * p.W;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    for (int j = 0; j < nx; ++j) {
      const long px = row + xidx[j];
      const float wv = xwt[j];
      if (F32) {                                                     // depth: one channel, DepthNorm first
        float v = __ldg(static_cast<const float*>(p.src) + px) / p.pre_div;
        v = fminf(fmaxf(v, p.lo), p.hi) / p.post_div;
        r0 = fmaf(wv, v, r0);
      } else {                                                       // uint8 HWC, 3 channels: ToTensor = x / 255
        const unsigned char* q = static_cast<const unsigned char*>(p.src) + 3 * px;
        r0 = fmaf(wv, lut[__ldg(q)], r0);
        r1 = fmaf(wv, lut[__ldg(q + 1)], r1);
        r2 = fmaf(wv, lut[__ldg(q + 2)], r2);
      }
    }
    const float wy = yw[ty * T + i];
    acc0 = fmaf(wy, r0, acc0);
    if (!F32) acc1 = fmaf(wy, r1, acc1), acc2 = fmaf(wy, r2, acc2);
  }
  if (F32) acc1 = acc2 = acc0;                                       // .unsqueeze(0).repeat(3, 1, 1)
  const long plane = static_cast<long>(p.S) * p.S, o = static_cast<long>(oy) * p.S + ox;
  p.dst[o] = (acc0 - p.mean[0]) / p.std_[0];                         // Normalize: sub_(mean).div_(std)
  p.dst[plane + o] = (acc1 - p.mean[1]) / p.std_[1];
  p.dst[2 * plane + o] = (acc2 - p.mean[2]) / p.std_[2];
}

}  // namespace missm

extern "C" int missm_image_preprocess(const missm_preproc_args* a, void* stream) {
  using namespace missm;
  MISSM_REQUIRE(a != nullptr && a->src != nullptr && a->dst != nullptr, "preprocess: null pointer");
  MISSM_REQUIRE(a->H > 0 && a->W > 0 && a->S > 0, "preprocess: bad sizes H=%d W=%d S=%d", a->H, a->W, a->S);
  MISSM_REQUIRE(a->pre_div != 0.f && a->post_div != 0.f, "preprocess: zero divisor");
  for (int c = 0; c < 3; ++c) MISSM_REQUIRE(a->std_[c] != 0.f, "preprocess: std[%d] = 0", c);
  PreprocParams p;
  p.src = a->src, p.dst = a->dst, p.src_f32 = a->src_f32, p.H = a->H, p.W = a->W, p.S = a->S;
  p.antialias = a->antialias ? 1 : 0;
  // torchvision Resize(S): the SHORTER side becomes S, the other int(S * long / short); CenterCrop(S):
  // top = int(round((RH - S) / 2.0)) -- round-half-even, as Python's round
  if (a->H <= a->W) {
    p.RH = a->S, p.RW = static_cast<int>(static_cast<long>(a->S) * a->W / a->H);
  } else {
    p.RW = a->S, p.RH = static_cast<int>(static_cast<long>(a->S) * a->H / a->W);
  }
  auto half_even = [](int d) { return (d % 2 == 0) ? d / 2 : ((d / 2) % 2 == 0 ? d / 2 : d / 2 + 1); };
  p.top = half_even(p.RH - a->S), p.left = half_even(p.RW - a->S);
  int taps = 4;
  if (p.antialias) {
    const float sy = static_cast<float>(a->H) / p.RH, sx = static_cast<float>(a->W) / p.RW;
    const float s = sy > sx ? sy : sx;
    taps = static_cast<int>(2.0f * (s >= 1.0f ? 2.0f * s : 2.0f)) + 3;
  }
  MISSM_REQUIRE(taps <= 512, "preprocess: down-scaling factor too large for the tap tables (%d taps)", taps);
  p.max_taps = taps;
  p.pre_div = a->pre_div, p.lo = a->clip_lo, p.hi = a->clip_hi, p.post_div = a->post_div;
  for (int c = 0; c < 3; ++c) p.mean[c] = a->mean[c], p.std_[c] = a->std_[c];
  const size_t smem = static_cast<size_t>(kTile) * taps * 16;        // 2 index + 2 weight tables
  const dim3 grid((a->S + kTile - 1) / kTile, (a->S + kTile - 1) / kTile);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->src_f32) {
    if (smem > 48 * 1024)
      MISSM_CHECK_CUDA(cudaFuncSetAttribute(image_preprocess_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    image_preprocess_kernel<true><<<grid, kTile * kTile, smem, st>>>(p); note_launch();
  } else {
    if (smem > 48 * 1024)
      MISSM_CHECK_CUDA(cudaFuncSetAttribute(image_preprocess_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    image_preprocess_kernel<false><<<grid, kTile * kTile, smem, st>>>(p); note_launch();
  }
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
