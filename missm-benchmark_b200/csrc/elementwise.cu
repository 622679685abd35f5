// HBM-bound helper kernels of the tower path: weight casts, bias-gradient column sums,
// patch extraction, embedding backward, pooling, L2-normalise.  All are coalesced, 128-bit
// vectorised where the layout allows, and sized in multiples of the SM count.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

// ---------------------------------------------------------------------------------------
// fp32 -> bf16 cast of a [rows, cols] matrix into a (possibly wider, zero-padded) destination
// ---------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, long ld_src,
                                     __nv_bfloat16* __restrict__ dst, long ld_dst, int rows,
                                     int cols, int cols_dst) {
  // one thread handles 4 consecutive destination columns
  const long groups_per_row = cols_dst / 4;
  const long total = static_cast<long>(rows) * groups_per_row;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / groups_per_row;
    const int c = static_cast<int>(i % groups_per_row) * 4;
    float v[4];
    if (c + 3 < cols && (ld_src & 3) == 0) {
      float4 f = *reinterpret_cast<const float4*>(src + r * ld_src + c);
      v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (c + j < cols) ? src[r * ld_src + c + j] : 0.f;
    }
    *reinterpret_cast<uint2*>(dst + r * ld_dst + c) =
        make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
}

// ---------------------------------------------------------------------------------------
// column sums of a bf16 [M, N] matrix (bias gradients): stage 1 -> partial[R][N]
// ---------------------------------------------------------------------------------------
constexpr int kColsumRows = 32;  // threadIdx.y
__global__ void __launch_bounds__(32 * kColsumRows)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long ldx, int M, int N,
                   float* __restrict__ partial, int rows_per_block) {
  // block = 32 lanes (x: 8 columns each -> 256 columns) x 32 row-threads
  const int col0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col0 < N) {
    for (int r = r0 + threadIdx.y; r < r1; r += kColsumRows) {
      uint4 q = *reinterpret_cast<const uint4*>(x + static_cast<long>(r) * ldx + col0);
      float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z),
             d = unpack_bf16x2(q.w);
      acc[0] += a.x, acc[1] += a.y, acc[2] += b.x, acc[3] += b.y;
      acc[4] += c.x, acc[5] += c.y, acc[6] += d.x, acc[7] += d.y;
    }
  }
  __shared__ float red[kColsumRows][32 * 8 + 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 256) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < kColsumRows; ++r) s += red[r][tid];
    const int c = blockIdx.x * 256 + tid;
    if (c < N) partial[static_cast<long>(blockIdx.y) * N + c] = s;
  }
}

__global__ void __launch_bounds__(256)
reduce_rows_kernel(const float* __restrict__ partial, int R, int N, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < N)
    for (int r = threadIdx.y; r < R; r += 8) s += partial[static_cast<long>(r) * N + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    out[c] = t;
  }
}

// ---------------------------------------------------------------------------------------
// patch extraction (im2col of a stride = kernel conv):  pixels f32 [B, C, H, W] ->
// patches bf16 [B * gh * gw, Kpad], column k = (c * ps + i) * ps + j  (Conv2d weight order)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void patch_store(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void patch_store(float* p, float v) { *p = v; }
template <typename OutT>
__global__ void patchify_kernel(const float* __restrict__ px, const int* __restrict__ sample_index,
                                OutT* __restrict__ out, int Bn, int C, int T, int H, int W,
                                int ps, int gh, int gw, int Kpad) {
  // one warp per (sample, channel, pixel row): reads W contiguous floats (coalesced)
  const int lane = threadIdx.x & 31;
  const long warp_global = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  // input is [B, C, T, H, W] (T = 1: plain images); output image index = b * T + t
  const long total = static_cast<long>(Bn) * T * C * gh * ps;  // pixel rows that belong to some patch
  for (long w = warp_global; w < total; w += nwarps) {
    const int i = static_cast<int>(w % ps);
    const int py = static_cast<int>((w / ps) % gh);
    const int c = static_cast<int>((w / (static_cast<long>(ps) * gh)) % C);
    const int t = static_cast<int>((w / (static_cast<long>(ps) * gh * C)) % T);
    const int b = static_cast<int>(w / (static_cast<long>(ps) * gh * C * T));
    const long src_b = sample_index ? sample_index[b] : b;
    const float* row = px + (((src_b * C + c) * T + t) * H + (py * ps + i)) * W;
    for (int xcol = lane; xcol < gw * ps; xcol += 32) {
      const int pxi = xcol / ps, j = xcol % ps;
      const long prow = ((static_cast<long>(b) * T + t) * gh + py) * gw + pxi;
      patch_store(out + prow * Kpad + (c * ps + i) * ps + j, row[xcol]);
    }
  }
}
__global__ void zero_pad_cols_kernel(__nv_bfloat16* __restrict__ out, long rows, int K, int Kpad) {
  const int pad = Kpad - K;
  const long total = rows * pad;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    out[(i / pad) * Kpad + K + (i % pad)] = __float2bfloat16(0.f);
}

// ---------------------------------------------------------------------------------------
// embeddings: CLS row of every sequence:  tok[b, 0, :] = cls + pos[0]
// ---------------------------------------------------------------------------------------
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos,
                                float* __restrict__ tok, int Bn, int ntok, int D) {
  const long total = static_cast<long>(Bn) * D;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long b = i / D;
    tok[b * ntok * D + d] = cls[d] + pos[d];
  }
}

// embedding backward: dpos[t, :] = sum_b dtok[b, t, :];  patch-row grads as bf16 [B*P, D]
__global__ void embed_bwd_kernel(const float* __restrict__ dtok, float* __restrict__ dpos,
                                 __nv_bfloat16* __restrict__ dpatch, int Bn, int ntok, int D) {
  // thread -> (token t, 4 columns); loops over the batch
  const int groups = D / 4;
  const long total = static_cast<long>(ntok) * groups;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / groups);
    const int c = static_cast<int>(i % groups) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < Bn; ++b) {
      const float4 g =
          *reinterpret_cast<const float4*>(dtok + (static_cast<long>(b) * ntok + t) * D + c);
      acc.x += g.x, acc.y += g.y, acc.z += g.z, acc.w += g.w;
      if (t > 0)
        *reinterpret_cast<uint2*>(dpatch + (static_cast<long>(b) * (ntok - 1) + t - 1) * D + c) =
            make_uint2(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w));
    }
    *reinterpret_cast<float4*>(dpos + static_cast<long>(t) * D + c) = acc;
  }
}

// ---------------------------------------------------------------------------------------
// frame mean (video pooling):  out[b, :] = mean_t in[b*T + t, :]   and its backward
// ---------------------------------------------------------------------------------------
__global__ void frame_mean_kernel(const float* __restrict__ in, void* __restrict__ out,
                                  int out_bf16, int Bn, int T, int D) {
  const long total = static_cast<long>(Bn) * D;
  const float inv = 1.0f / T;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long b = i / D;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += in[(b * T + t) * D + d];
    if (out_bf16)
      reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16(s * inv);
    else
      reinterpret_cast<float*>(out)[i] = s * inv;
  }
}
__global__ void frame_mean_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din,
                                      int Bn, int T, int D) {
  const long total = static_cast<long>(Bn) * T * D;
  const float inv = 1.0f / T;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long b = i / (static_cast<long>(T) * D);
    din[i] = dout[b * D + d] * inv;
  }
}

// ---------------------------------------------------------------------------------------
// y = x / ||x||_2 * scale      (languagebind/__init__.py:80-83), one warp per row
// ---------------------------------------------------------------------------------------
__global__ void l2norm_scale_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                        float* __restrict__ inv_norm, float scale, int Bn, int P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Bn) return;
  const float* xr = x + static_cast<long>(warp) * P;
  float s = 0.f;
  for (int c = lane; c < P; c += 32) s += xr[c] * xr[c];
  const float inv = 1.0f / sqrtf(warp_sum(s));
  if (lane == 0 && inv_norm) inv_norm[warp] = inv;
  for (int c = lane; c < P; c += 32) y[static_cast<long>(warp) * P + c] = xr[c] * inv * scale;
}
// dx = scale * inv * (dy - yhat * <dy, yhat>),  yhat = x * inv ; dx written as bf16 (GEMM operand)
__global__ void l2norm_scale_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                        const float* __restrict__ inv_norm, float scale,
                                        void* __restrict__ dx, int dx_bf16, int Bn, int P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Bn) return;
  const float inv = inv_norm[warp];
  const float* xr = x + static_cast<long>(warp) * P;
  const float* dr = dy + static_cast<long>(warp) * P;
  float dot = 0.f;
  for (int c = lane; c < P; c += 32) dot += dr[c] * xr[c] * inv;
  dot = warp_sum(dot);
  for (int c = lane; c < P; c += 32) {
    const float v = scale * inv * (dr[c] - xr[c] * inv * dot);
    if (dx_bf16)
      reinterpret_cast<__nv_bfloat16*>(dx)[static_cast<long>(warp) * P + c] = __float2bfloat16(v);
    else
      reinterpret_cast<float*>(dx)[static_cast<long>(warp) * P + c] = v;
  }
}

// ---------------------------------------------------------------------------------------
// text tower embeddings (transformers 4.3x CLIPTextEmbeddings, used at modeling_image.py:463,494)
// ---------------------------------------------------------------------------------------
__global__ void text_embed_fwd_kernel(const int64_t* __restrict__ ids, const int* __restrict__ sample_index,
                                      const float* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                      float* __restrict__ out, int Bn, int L, int D) {
  const int groups = D / 4;
  const long total = static_cast<long>(Bn) * L * groups;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % groups) * 4;
    const long row = i / groups;
    const int l = static_cast<int>(row % L);
    const long b = row / L;
    const long sb = sample_index ? sample_index[b] : b;
    const int64_t id = ids[sb * L + l];
    const float4 a = *reinterpret_cast<const float4*>(tok_emb + id * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pos_emb + static_cast<long>(l) * D + c);
    *reinterpret_cast<float4*>(out + row * D + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}
// dtok_emb[id] += dx[row] (atomic scatter-add), dpos[l] = sum_b dx[b, l]
__global__ void text_embed_bwd_kernel(const int64_t* __restrict__ ids, const int* __restrict__ sample_index,
                                      const float* __restrict__ dx, float* __restrict__ dtok_emb,
                                      float* __restrict__ dpos, int Bn, int L, int D) {
  const long total = static_cast<long>(L) * D;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const int l = static_cast<int>(i / D);
    float s = 0.f;
    for (int b = 0; b < Bn; ++b) {
      const float g = dx[(static_cast<long>(b) * L + l) * D + d];
      s += g;
      const long sb = sample_index ? sample_index[b] : b;
      atomicAdd(dtok_emb + ids[sb * L + l] * D + d, g);
    }
    dpos[i] = s;
  }
}
// first index of the maximum id per row (torch.argmax tie rule on the EOT token, :519-522)
__global__ void argmax_rows_kernel(const int64_t* __restrict__ ids, const int* __restrict__ sample_index,
                                   int* __restrict__ out_rows, int Bn, int L) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bn) return;
  const long sb = sample_index ? sample_index[b] : b;
  const int64_t* r = ids + sb * L;
  int best = 0;
  int bv = static_cast<int>(r[0]);  // reference casts to int32 before argmax
  for (int l = 1; l < L; ++l) {
    const int v = static_cast<int>(r[l]);
    if (v > bv) bv = v, best = l;
  }
  out_rows[b] = b * L + best;  // row in the compacted [Bn*L, D] token matrix
}

// out[g, :] = sum over rows r with (r / div) % period == g of x[r, :]   (x f32 [M, D], D % 4 == 0)
// (gradient of the temporal embedding, modeling_image.py:113).  HBM-bound: one read of x.  Rows of group g come in
// runs of `div` consecutive rows, one run per `period * div` rows; blockIdx.z walks (run, slice of a run), every
// thread sums 4 columns over its rows (4 loads in flight) and adds them to the pre-zeroed output.
__global__ void __launch_bounds__(256)
colsum_grouped_kernel(const float* __restrict__ x, int M, int D, int period, int div, int kGroupedSub,
                      float* __restrict__ out) {
  const int g = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const long span = static_cast<long>(period) * div;
  const int run = blockIdx.z / kGroupedSub, sub = blockIdx.z % kGroupedSub;
  const int per = (div + kGroupedSub - 1) / kGroupedSub;
  const long r_begin = run * span + static_cast<long>(g) * div + static_cast<long>(sub) * per;
  long r_end = r_begin + per;
  const long run_end = run * span + static_cast<long>(g + 1) * div;
  if (r_end > run_end) r_end = run_end;
  if (r_end > M) r_end = M;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  long r = r_begin;
  for (; r + 3 < r_end; r += 4) {
    const float4 v0 = *reinterpret_cast<const float4*>(x + r * D + c);
    const float4 v1 = *reinterpret_cast<const float4*>(x + (r + 1) * D + c);
    const float4 v2 = *reinterpret_cast<const float4*>(x + (r + 2) * D + c);
    const float4 v3 = *reinterpret_cast<const float4*>(x + (r + 3) * D + c);
    a0.x += v0.x, a0.y += v0.y, a0.z += v0.z, a0.w += v0.w;
    a1.x += v1.x, a1.y += v1.y, a1.z += v1.z, a1.w += v1.w;
    a2.x += v2.x, a2.y += v2.y, a2.z += v2.z, a2.w += v2.w;
    a3.x += v3.x, a3.y += v3.y, a3.z += v3.z, a3.w += v3.w;
  }
  for (; r < r_end; ++r) {
    const float4 v0 = *reinterpret_cast<const float4*>(x + r * D + c);
    a0.x += v0.x, a0.y += v0.y, a0.z += v0.z, a0.w += v0.w;
  }
  if (r_begin >= r_end) return;
  float* o = out + static_cast<long>(g) * D + c;
  atomicAdd(o, (a0.x + a1.x) + (a2.x + a3.x));
  atomicAdd(o + 1, (a0.y + a1.y) + (a2.y + a3.y));
  atomicAdd(o + 2, (a0.z + a1.z) + (a2.z + a3.z));
  atomicAdd(o + 3, (a0.w + a1.w) + (a2.w + a3.w));
}

static inline int grid_for(long total, int threads) {
  long b = (total + threads - 1) / threads;
  long cap = 16L * kNumSMs;
  if (b < 1) b = 1;
  return static_cast<int>(b < cap ? b : cap);
}

}  // namespace missm

using namespace missm;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int missm_cast_f32_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst,
                                   int32_t rows, int32_t cols, int32_t cols_dst, void* stream) {
  if (rows == 0) return 0;
  MISSM_REQUIRE(cols_dst % 4 == 0 && ld_dst % 4 == 0 && cols_dst >= cols, "cast: bad dst cols %d", cols_dst);
  const long total = static_cast<long>(rows) * (cols_dst / 4);
  cast_f32_bf16_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(
      src, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst, rows, cols, cols_dst); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- small-footprint variant (default): 128 threads, <= 32 registers, no shared memory, so that a CTA fits into the
// registers / thread slots / shared memory the persistent tcgen05 kernels leave free on every SM (kCoResidentRegs,
// missm_common.cuh).  The pass is HBM-bound and the GEMMs of the other towers' streams are tensor-bound: co-resident,
// it runs on bandwidth nobody else is using instead of waiting for a gap between two persistent kernels.
// One thread = 8 columns (one 16-byte load per row), one CTA = 1024 columns x rows_per_block rows, no cross-thread
// reduction at all: partial[blockIdx.y][col] is written straight from the accumulators.
__global__ void __launch_bounds__(128, 16)
colsum_bf16_small_kernel(const __nv_bfloat16* __restrict__ x, long ldx, int M, int N, float* __restrict__ partial,
                         int rows_per_block) {
  const int col0 = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (col0 >= N) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* p = x + static_cast<long>(r0) * ldx + col0;
  auto add = [&](const uint4& q) {
    const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
    acc[0] += a.x, acc[1] += a.y, acc[2] += b.x, acc[3] += b.y;
    acc[4] += c.x, acc[5] += c.y, acc[6] += d.x, acc[7] += d.y;
  };
  int r = r0;
  for (; r + 4 <= r1; r += 4, p += 4 * ldx) {      // four independent 16-byte loads in flight per thread
    const uint4 q0 = *reinterpret_cast<const uint4*>(p);
    const uint4 q1 = *reinterpret_cast<const uint4*>(p + ldx);
    const uint4 q2 = *reinterpret_cast<const uint4*>(p + 2 * ldx);
    const uint4 q3 = *reinterpret_cast<const uint4*>(p + 3 * ldx);
    add(q0), add(q1), add(q2), add(q3);
  }
  for (; r < r1; ++r, p += ldx) add(*reinterpret_cast<const uint4*>(p));
  float4* dst = reinterpret_cast<float4*>(partial + static_cast<long>(blockIdx.y) * N + col0);
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}
// second stage, same footprint: a warp owns 8 columns x 4 row groups (32-byte row segments, the partials sit in L2),
// two shuffles fold the row groups; fixed summation order
__global__ void __launch_bounds__(128, 16)
reduce_rows_small_kernel(const float* __restrict__ partial, int R, int N, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + warp * 8 + (lane & 7);
  const int g = lane >> 3;
  float s0 = 0.f, s1 = 0.f;
  if (c < N) {
    int r = g;
    for (; r + 4 < R; r += 8) s0 += partial[static_cast<long>(r) * N + c], s1 += partial[static_cast<long>(r + 4) * N + c];
    if (r < R) s0 += partial[static_cast<long>(r) * N + c];
  }
  float s = s0 + s1;
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  if (g == 0 && c < N) out[c] = s;
}

static bool colsum_small() {
  static const bool on = getenv("MISSM_COLSUM_SMALL") == nullptr || atoi(getenv("MISSM_COLSUM_SMALL")) != 0;   // A/B switch
  return on;
}
extern "C" int missm_colsum_num_partials(int32_t M) {
  if (colsum_small()) {
    int r = (M + 63) / 64;
    if (r < 1) r = 1;
    return r < 256 ? r : 256;
  }
  int r = (M + 255) / 256;
  if (r < 1) r = 1;
  return r < 64 ? r : 64;
}
// partial: workspace [missm_colsum_num_partials(M), N] floats
extern "C" int missm_colsum_bf16(const void* x, int64_t ldx, int32_t M, int32_t N, float* partial,
                                 float* out, void* stream) {
  MISSM_REQUIRE(N % 8 == 0 && ldx % 8 == 0, "colsum: N=%d ldx=%ld must be multiples of 8", N, (long)ldx);
  if (M == 0) {
    MISSM_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, ST(stream)));
    return 0;
  }
  const int R = missm_colsum_num_partials(M);
  const int rows_per_block = (M + R - 1) / R;
  if (colsum_small()) {
    dim3 grid((N + 1023) / 1024, R);
    colsum_bf16_small_kernel<<<grid, 128, 0, ST(stream)>>>(static_cast<const __nv_bfloat16*>(x), ldx, M, N, partial,
                                                          rows_per_block); note_launch();
    reduce_rows_small_kernel<<<(N + 31) / 32, 128, 0, ST(stream)>>>(partial, R, N, out); note_launch();
    MISSM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid((N + 255) / 256, R), block(32, kColsumRows);
  colsum_bf16_kernel<<<grid, block, 0, ST(stream)>>>(static_cast<const __nv_bfloat16*>(x), ldx, M, N,
                                                    partial, rows_per_block); note_launch();
  reduce_rows_kernel<<<(N + 31) / 32, dim3(32, 8), 0, ST(stream)>>>(partial, R, N, out); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_patchify(const float* pixels, const int32_t* sample_index, void* patches,
                              int32_t Bn, int32_t C, int32_t T, int32_t H, int32_t W, int32_t ps,
                              int32_t Kpad, void* stream) {
  if (Bn == 0) return 0;
  const int gh = H / ps, gw = W / ps, K = C * ps * ps;
  MISSM_REQUIRE(Kpad >= K && Kpad % 8 == 0 && T >= 1, "patchify: Kpad=%d K=%d T=%d", Kpad, K, T);
  const long rows = static_cast<long>(Bn) * T * gh * gw;
  if (Kpad > K) {
    zero_pad_cols_kernel<<<grid_for(rows * (Kpad - K), 256), 256, 0, ST(stream)>>>(
        static_cast<__nv_bfloat16*>(patches), rows, K, Kpad); note_launch();
  }
  const long warps = static_cast<long>(Bn) * T * C * gh * ps;
  patchify_kernel<__nv_bfloat16><<<grid_for(warps * 32, 256), 256, 0, ST(stream)>>>(
      pixels, sample_index, static_cast<__nv_bfloat16*>(patches), Bn, C, T, H, W, ps, gh, gw, Kpad); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

/* fp32 verification mode: patches f32 [rows, Kpad]; columns K = C * ps * ps .. Kpad-1 are NOT written
   (the caller zero-fills the buffer) */
extern "C" int missm_patchify_f32(const float* pixels, const int32_t* sample_index, float* patches,
                                  int32_t Bn, int32_t C, int32_t T, int32_t H, int32_t W, int32_t ps, int32_t Kpad,
                                  void* stream) {
  if (Bn == 0) return 0;
  const int gh = H / ps, gw = W / ps, K = C * ps * ps;
  MISSM_REQUIRE(T >= 1 && Kpad >= K, "patchify_f32: T=%d Kpad=%d K=%d", T, Kpad, K);
  const long warps = static_cast<long>(Bn) * T * C * gh * ps;
  patchify_kernel<float><<<grid_for(warps * 32, 256), 256, 0, ST(stream)>>>(pixels, sample_index, patches, Bn, C, T, H,
                                                                            W, ps, gh, gw, Kpad); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_cls_rows(const float* cls, const float* pos, float* tok, int32_t Bn,
                              int32_t ntok, int32_t D, void* stream) {
  if (Bn == 0) return 0;
  cls_rows_kernel<<<grid_for(static_cast<long>(Bn) * D, 256), 256, 0, ST(stream)>>>(cls, pos, tok, Bn, ntok, D); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_embed_bwd(const float* dtok, float* dpos, void* dpatch_bf16, int32_t Bn,
                               int32_t ntok, int32_t D, void* stream) {
  MISSM_REQUIRE(D % 4 == 0, "embed_bwd: D=%d", D);
  embed_bwd_kernel<<<grid_for(static_cast<long>(ntok) * (D / 4), 128), 128, 0, ST(stream)>>>(
      dtok, dpos, static_cast<__nv_bfloat16*>(dpatch_bf16), Bn, ntok, D); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_frame_mean(const float* in, void* out, int32_t out_bf16, int32_t Bn, int32_t T,
                                int32_t D, void* stream) {
  if (Bn == 0) return 0;
  frame_mean_kernel<<<grid_for(static_cast<long>(Bn) * D, 256), 256, 0, ST(stream)>>>(in, out, out_bf16, Bn, T, D); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int missm_frame_mean_bwd(const float* dout, float* din, int32_t Bn, int32_t T, int32_t D,
                                    void* stream) {
  if (Bn == 0) return 0;
  frame_mean_bwd_kernel<<<grid_for(static_cast<long>(Bn) * T * D, 256), 256, 0, ST(stream)>>>(dout, din, Bn, T, D); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_l2norm_scale_fwd(const float* x, float* y, float* inv_norm, float scale,
                                      int32_t Bn, int32_t P, void* stream) {
  if (Bn == 0) return 0;
  l2norm_scale_fwd_kernel<<<(Bn * 32 + 127) / 128, 128, 0, ST(stream)>>>(x, y, inv_norm, scale, Bn, P); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int missm_l2norm_scale_bwd(const float* dy, const float* x, const float* inv_norm,
                                      float scale, void* dx, int32_t dx_bf16, int32_t Bn, int32_t P,
                                      void* stream) {
  if (Bn == 0) return 0;
  l2norm_scale_bwd_kernel<<<(Bn * 32 + 127) / 128, 128, 0, ST(stream)>>>(dy, x, inv_norm, scale, dx, dx_bf16, Bn, P); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_text_embed_fwd(const int64_t* ids, const int32_t* sample_index,
                                    const float* tok_emb, const float* pos_emb, float* out,
                                    int32_t Bn, int32_t L, int32_t D, void* stream) {
  if (Bn == 0) return 0;
  MISSM_REQUIRE(D % 4 == 0, "text_embed: D=%d", D);
  text_embed_fwd_kernel<<<grid_for(static_cast<long>(Bn) * L * (D / 4), 256), 256, 0, ST(stream)>>>(
      ids, sample_index, tok_emb, pos_emb, out, Bn, L, D); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// dtok_emb must be zero-initialised by the caller ([vocab, D]); dpos is [L, D] (overwritten)
extern "C" int missm_text_embed_bwd(const int64_t* ids, const int32_t* sample_index, const float* dx,
                                    float* dtok_emb, float* dpos, int32_t Bn, int32_t L, int32_t D,
                                    void* stream) {
  text_embed_bwd_kernel<<<grid_for(static_cast<long>(L) * D, 128), 128, 0, ST(stream)>>>(
      ids, sample_index, dx, dtok_emb, dpos, Bn, L, D); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int missm_argmax_rows(const int64_t* ids, const int32_t* sample_index, int32_t* out_rows,
                                 int32_t Bn, int32_t L, void* stream) {
  if (Bn == 0) return 0;
  argmax_rows_kernel<<<(Bn + 127) / 128, 128, 0, ST(stream)>>>(ids, sample_index, out_rows, Bn, L); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_colsum_grouped_f32(const float* x, int32_t M, int32_t D, int32_t period, int32_t div,
                                        float* out, void* stream) {
  MISSM_REQUIRE(period > 0 && div > 0 && D % 4 == 0, "colsum_grouped: period=%d div=%d D=%d", period, div, D);
  MISSM_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(period) * D, ST(stream)));
  if (M == 0) return 0;
  if (period == 1) div = M;                       // plain column sums: one run of all rows
  const long span = static_cast<long>(period) * div;
  const int runs = static_cast<int>((M + span - 1) / span);
  const int gx = (D / 4 + 255) / 256;
  // slices per run: enough blocks for ~8 per SM, at least 16 rows each
  long sub = (8L * kNumSMs + static_cast<long>(gx) * period * runs - 1) / (static_cast<long>(gx) * period * runs);
  if (sub > (div + 15) / 16) sub = (div + 15) / 16;
  if (sub < 1) sub = 1;
  MISSM_REQUIRE(runs * sub <= 65535, "colsum_grouped: too many runs (%d)", runs);
  dim3 grid(gx, period, static_cast<unsigned>(runs * sub));
  colsum_grouped_kernel<<<grid, 256, 0, ST(stream)>>>(x, M, D, period, div, static_cast<int>(sub), out); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_copy_f32(const float* src, float* dst, int64_t n, void* stream) {
  if (n == 0) return 0;
  MISSM_CHECK_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, ST(stream)));
  return 0;
}
