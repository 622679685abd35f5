// LayerNorm forward / backward over rows of an fp32 residual stream (HBM-bound).
//
// Replaces nn.LayerNorm at languagebind/image/modeling_image.py:120,132,139,149 (layer norms of
// CLIPEncoderLayer), :649 (pre_layrnorm), :660 (post_layernorm on CLS rows), :514
// (final_layer_norm; applied only to the gathered EOT rows -- LN is row-wise so gathering first
// is exact), and src/model/baseline.py:61 (fusion norm).
//
// One warp per row, the whole row lives in registers (D = 128*VEC, VEC <= 12), two-pass
// mean/variance (biased, as torch), 128-bit loads/stores.
//   fwd bytes/row : read 4D (+ optional gather), write 2D (bf16 out) or 4D (f32 out) + 8
//   bwd bytes/row : read 2D|4D (dy) + 4D (x) + 4D (dres, optional), write 4D (+2D bf16 copy)
#include <cstdlib>

#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

constexpr int kLnWarps = 8;

template <int VEC, bool OUT_BF16>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const float* __restrict__ x, long ldx, const int* __restrict__ row_index,
                     const float* __restrict__ add_rows, int add_period, int add_div,
                     float* __restrict__ x_out,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     void* __restrict__ y, long ldy, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, float eps) {
  constexpr int D = VEC * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * kLnWarps + warp; row < M; row += gridDim.x * kLnWarps) {
    const long src_row = row_index ? row_index[row] : row;
    const float4* xr = reinterpret_cast<const float4*>(x + src_row * ldx);
    float4 v[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = xr[lane + 32 * i];
    if (add_rows != nullptr) {
      // x <- x + add_rows[(row / add_div) % add_period]   (temporal embedding, written back)
      const int t = (row / add_div) % add_period;
      const float4* ar = reinterpret_cast<const float4*>(add_rows + static_cast<long>(t) * D);
      float4* xo = reinterpret_cast<float4*>(x_out + src_row * ldx);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float4 a = __ldg(ar + lane + 32 * i);
        v[i].x += a.x, v[i].y += a.y, v[i].z += a.z, v[i].w += a.w;
        xo[lane + 32 * i] = v[i];
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if constexpr (OUT_BF16) {
        uint2 pk = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * ldy)[lane + 32 * i] = pk;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * ldy)[lane + 32 * i] = o;
      }
    }
  }
}

// dx[row] = (dres ? dres[row] : 0) + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy*gamma
// partial[blockIdx][0][:] = sum_rows dy * xhat ; partial[blockIdx][1][:] = sum_rows dy ;
// partial[blockIdx][2][:] = sum_rows dx
template <int VEC, bool DY_BF16>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_kernel(const void* __restrict__ dy, long lddy, const float* __restrict__ x, long ldx,
                     const int* __restrict__ row_index, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ dres, float* __restrict__ dx,
                     __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ partial, int M) {
  constexpr int D = VEC * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 dg[VEC], db[VEC], ds[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) dg[i] = db[i] = ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int row = blockIdx.x * kLnWarps + warp; row < M; row += gridDim.x * kLnWarps) {
    const long xrow = row_index ? row_index[row] : row;
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + xrow * ldx);
    float4 xh[VEC], g[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 d;
      if constexpr (DY_BF16) {
        uint2 pk = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) +
                                                  row * lddy)[lane + 32 * i];
        float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
        d = make_float4(a.x, a.y, b.x, b.y);
      } else {
        d = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * lddy)[lane + 32 * i];
      }
      const float4 xv = xr[lane + 32 * i];
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      dg[i].x += d.x * xh[i].x, dg[i].y += d.y * xh[i].y, dg[i].z += d.z * xh[i].z, dg[i].w += d.w * xh[i].w;
      db[i].x += d.x, db[i].y += d.y, db[i].z += d.z, db[i].w += d.w;
    }
    const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 o;
      o.x = rs * (g[i].x - c1 - xh[i].x * c2);
      o.y = rs * (g[i].y - c1 - xh[i].y * c2);
      o.z = rs * (g[i].z - c1 - xh[i].z * c2);
      o.w = rs * (g[i].w - c1 - xh[i].w * c2);
      if (dres != nullptr) {
        const float4 r = reinterpret_cast<const float4*>(dres + xrow * ldx)[lane + 32 * i];
        o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
      }
      reinterpret_cast<float4*>(dx + xrow * ldx)[lane + 32 * i] = o;
      ds[i].x += o.x, ds[i].y += o.y, ds[i].z += o.z, ds[i].w += o.w;
      if (dx_bf16 != nullptr)
        reinterpret_cast<uint2*>(dx_bf16 + xrow * ldx)[lane + 32 * i] =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }

  // block reduction of the per-warp column sums through shared memory
  __shared__ float4 red[kLnWarps][32];
  float* pg = partial + static_cast<long>(blockIdx.x) * 3 * D;
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
#pragma unroll 1
    for (int i = 0; i < VEC; ++i) {
      red[warp][lane] = which == 0 ? dg[i] : (which == 1 ? db[i] : ds[i]);
      __syncthreads();
      if (warp == 0) {
        float4 a = red[0][lane];
#pragma unroll
        for (int w = 1; w < kLnWarps; ++w) {
          float4 b = red[w][lane];
          a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        reinterpret_cast<float4*>(pg + which * D)[lane + 32 * i] = a;
      }
      __syncthreads();
    }
  }
}

// The same for wide rows (D >= 768): TWO warps per row, 16 warps per CTA.  With one warp per row a thread carries
// 3 x D/32 column accumulators plus the row itself (~200 registers for D = 1024): 8 warps per SM, too few loads
// in flight to cover the HBM latency (measured 4.7 TB/s).  Halving the columns per warp halves the registers, so
// a 512-thread CTA fits (<= 128 registers) and twice the bytes are in flight.  The two row sums are exchanged
// through shared memory and a 64-thread named barrier per row pair (double-buffered by iteration parity).
constexpr int kLn2Warps = 16;

template <int HV, bool DY_BF16>
__global__ void __launch_bounds__(kLn2Warps * 32, 1)
layernorm_bwd2_kernel(const void* __restrict__ dy, long lddy, const float* __restrict__ x, long ldx,
                      const int* __restrict__ row_index, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                      const float* __restrict__ dres, float* __restrict__ dx,
                      __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ partial, int M) {
  constexpr int D = 2 * HV * 128;
  constexpr int PAIRS = kLn2Warps / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, half = warp & 1;
  const int c0 = half * HV * 32 + lane;          // float4 index of this lane's first column block
  __shared__ float2 xch[2][PAIRS][2];
  __shared__ float4 red[kLn2Warps][32];
  float4 dg[HV], db[HV], ds[HV];
#pragma unroll
  for (int i = 0; i < HV; ++i) dg[i] = db[i] = ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  int par = 0;
  for (int row = blockIdx.x * PAIRS + pair; row < M; row += gridDim.x * PAIRS, par ^= 1) {
    const long xrow = row_index ? row_index[row] : row;
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + xrow * ldx);
    float4 xh[HV], g[HV], r[HV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < HV; ++i) {     // all loads of the row first
      xh[i] = xr[c0 + 32 * i];
      if (dres != nullptr) r[i] = reinterpret_cast<const float4*>(dres + xrow * ldx)[c0 + 32 * i];
      if constexpr (DY_BF16) {
        const uint2 pk = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + row * lddy)[c0 + 32 * i];
        const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
        g[i] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        g[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + row * lddy)[c0 + 32 * i];
      }
    }
#pragma unroll
    for (int i = 0; i < HV; ++i) {
      const float4 d = g[i];
      const float4 gmi = __ldg(reinterpret_cast<const float4*>(gamma) + c0 + 32 * i);
      xh[i] = make_float4((xh[i].x - mu) * rs, (xh[i].y - mu) * rs, (xh[i].z - mu) * rs, (xh[i].w - mu) * rs);
      g[i] = make_float4(d.x * gmi.x, d.y * gmi.y, d.z * gmi.z, d.w * gmi.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      dg[i].x += d.x * xh[i].x, dg[i].y += d.y * xh[i].y, dg[i].z += d.z * xh[i].z, dg[i].w += d.w * xh[i].w;
      db[i].x += d.x, db[i].y += d.y, db[i].z += d.z, db[i].w += d.w;
    }
    s1 = warp_sum(s1), s2 = warp_sum(s2);
    if (lane == 0) xch[par][pair][half] = make_float2(s1, s2);
    asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
    const float2 other = xch[par][pair][half ^ 1];
    const float c1 = (s1 + other.x) * (1.0f / D), c2 = (s2 + other.y) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < HV; ++i) {
      float4 o;
      o.x = rs * (g[i].x - c1 - xh[i].x * c2);
      o.y = rs * (g[i].y - c1 - xh[i].y * c2);
      o.z = rs * (g[i].z - c1 - xh[i].z * c2);
      o.w = rs * (g[i].w - c1 - xh[i].w * c2);
      if (dres != nullptr) o.x += r[i].x, o.y += r[i].y, o.z += r[i].z, o.w += r[i].w;
      reinterpret_cast<float4*>(dx + xrow * ldx)[c0 + 32 * i] = o;
      ds[i].x += o.x, ds[i].y += o.y, ds[i].z += o.z, ds[i].w += o.w;
      if (dx_bf16 != nullptr)
        reinterpret_cast<uint2*>(dx_bf16 + xrow * ldx)[c0 + 32 * i] =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }

  // block reduction over the row pairs, per column half
  float* pg = partial + static_cast<long>(blockIdx.x) * 3 * D;
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
#pragma unroll 1
    for (int i = 0; i < HV; ++i) {
      red[warp][lane] = which == 0 ? dg[i] : (which == 1 ? db[i] : ds[i]);
      __syncthreads();
      if (warp < 2) {
        float4 a = red[warp][lane];
#pragma unroll
        for (int w = 1; w < PAIRS; ++w) {
          const float4 b = red[2 * w + warp][lane];
          a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        reinterpret_cast<float4*>(pg + which * D)[c0 + 32 * i] = a;
      }
      __syncthreads();
    }
  }
}

// out[c] = sum_r partial[r][c]    (deterministic second stage of column reductions)
// block = 32 columns x 8 row-threads: coalesced 128 B row segments, fixed summation order
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int R, long stride,
                       float* __restrict__ out, int n, float scale) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < n)
    for (int r = threadIdx.y; r < R; r += 8) s += partial[r * stride + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    out[c] = t * scale;
  }
}

template <bool OUT_BF16>
static int launch_ln_fwd(int vec, int grid, cudaStream_t st, const float* x, long ldx,
                         const int* ridx, const float* add_rows, int add_period, int add_div,
                         float* x_out, const float* g, const float* b, void* y, long ldy,
                         float* mean, float* rstd, int M, float eps) {
#define LN_CASE(V)                                                                              \
  case V:                                                                                       \
    layernorm_fwd_kernel<V, OUT_BF16><<<grid, kLnWarps * 32, 0, st>>>(                          \
        x, ldx, ridx, add_rows, add_period, add_div, x_out, g, b, y, ldy, mean, rstd, M, eps); note_launch();  \
    break;
  switch (vec) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(8) LN_CASE(10)
    LN_CASE(12)
    default:
      MISSM_REQUIRE(false, "layernorm: unsupported width %d", vec * 128);
  }
#undef LN_CASE
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <bool DY_BF16>
static int launch_ln_bwd(int vec, int grid, cudaStream_t st, const void* dy, long lddy,
                         const float* x, long ldx, const int* ridx, const float* mean,
                         const float* rstd, const float* gamma, const float* dres, float* dx,
                         __nv_bfloat16* dxb, float* partial, int M) {
#define LN_CASE(V)                                                                         \
  case V:                                                                                  \
    layernorm_bwd_kernel<V, DY_BF16><<<grid, kLnWarps * 32, 0, st>>>(                      \
        dy, lddy, x, ldx, ridx, mean, rstd, gamma, dres, dx, dxb, partial, M); note_launch();             \
    break;
#define LN2_CASE(V)                                                                        \
  case V:                                                                                  \
    layernorm_bwd2_kernel<V / 2, DY_BF16><<<grid, kLn2Warps * 32, 0, st>>>(                \
        dy, lddy, x, ldx, ridx, mean, rstd, gamma, dres, dx, dxb, partial, M); note_launch();             \
    break;
  static const bool one_warp_rows = getenv("MISSM_LN_BWD_1WARP") != nullptr;   // A/B switch
  if ((vec == 6 || vec == 8) && !one_warp_rows) {
    switch (vec) {
      LN2_CASE(6) LN2_CASE(8)
      default:
        MISSM_REQUIRE(false, "layernorm: unsupported width %d", vec * 128);
    }
  } else {
    switch (vec) {
      LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(8) LN_CASE(10)
      LN_CASE(12)
      default:
        MISSM_REQUIRE(false, "layernorm: unsupported width %d", vec * 128);
    }
  }
#undef LN_CASE
#undef LN2_CASE
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace missm

using namespace missm;

extern "C" int missm_ln_bwd_num_partials(int M) {
  int blocks = (M + kLnWarps - 1) / kLnWarps;
  int cap = kNumSMs;
  return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

extern "C" int missm_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_index,
                                   const float* add_rows, int32_t add_period, int32_t add_div,
                                   float* x_out, const float* gamma, const float* beta, void* y,
                                   int64_t ldy, int32_t y_bf16, float* mean, float* rstd, int32_t M,
                                   int32_t D, float eps, void* stream) {
  if (M == 0) return 0;
  MISSM_REQUIRE(D % 128 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "layernorm: D=%d ldx=%ld ldy=%ld", D,
                (long)ldx, (long)ldy);
  MISSM_REQUIRE(add_rows == nullptr || (x_out != nullptr && add_period > 0 && add_div > 0),
                "layernorm: add_rows needs x_out/add_period/add_div");
  const int blocks = (M + kLnWarps - 1) / kLnWarps;
  const int grid = blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (y_bf16)
    return launch_ln_fwd<true>(D / 128, grid, st, x, ldx, row_index, add_rows, add_period, add_div,
                               x_out, gamma, beta, y, ldy, mean, rstd, M, eps);
  return launch_ln_fwd<false>(D / 128, grid, st, x, ldx, row_index, add_rows, add_period, add_div,
                              x_out, gamma, beta, y, ldy, mean, rstd, M, eps);
}

// partial: workspace of missm_ln_bwd_num_partials(M) * 3 * D floats; dgamma/dbeta/dx_colsum: [D].
extern "C" int missm_layernorm_bwd(const void* dy, int64_t lddy, int32_t dy_bf16, const float* x,
                                   int64_t ldx, const int32_t* row_index, const float* mean,
                                   const float* rstd, const float* gamma, const float* dres,
                                   float* dx, void* dx_bf16, float* partial, float* dgamma,
                                   float* dbeta, float* dx_colsum, int32_t M, int32_t D, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MISSM_REQUIRE(D % 128 == 0, "layernorm: D=%d", D);
  if (M == 0) {
    MISSM_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * D, st));
    MISSM_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * D, st));
    if (dx_colsum) MISSM_CHECK_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * D, st));
    return 0;
  }
  const int grid = missm_ln_bwd_num_partials(M);
  int rc;
  if (dy_bf16)
    rc = launch_ln_bwd<true>(D / 128, grid, st, dy, lddy, x, ldx, row_index, mean, rstd, gamma,
                             dres, dx, static_cast<__nv_bfloat16*>(dx_bf16), partial, M);
  else
    rc = launch_ln_bwd<false>(D / 128, grid, st, dy, lddy, x, ldx, row_index, mean, rstd, gamma,
                              dres, dx, static_cast<__nv_bfloat16*>(dx_bf16), partial, M);
  if (rc) return rc;
  if (dbeta == dgamma + D && (dx_colsum == nullptr || dx_colsum == dbeta + D)) {  // contiguous output: one launch
    const int n = (dx_colsum ? 3 : 2) * D;
    reduce_partials_kernel<<<(n + 31) / 32, dim3(32, 8), 0, st>>>(partial, grid, 3L * D, dgamma, n, 1.f); note_launch();
  } else {
    reduce_partials_kernel<<<(D + 31) / 32, dim3(32, 8), 0, st>>>(partial, grid, 3L * D, dgamma, D, 1.f); note_launch();
    reduce_partials_kernel<<<(D + 31) / 32, dim3(32, 8), 0, st>>>(partial + D, grid, 3L * D, dbeta, D, 1.f); note_launch();
    if (dx_colsum)
      reduce_partials_kernel<<<(D + 31) / 32, dim3(32, 8), 0, st>>>(partial + 2 * D, grid, 3L * D, dx_colsum, D, 1.f); note_launch();
  }
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_reduce_partials(const float* partial, int32_t R, int64_t stride, float* out,
                                     int32_t n, float scale, void* stream) {
  reduce_partials_kernel<<<(n + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      partial, R, stride, out, n, scale); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
