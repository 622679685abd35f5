// fp32 VERIFICATION mode (MISSM_PRECISION=fp32): the same path with fp32-grade arithmetic, to check the bf16
// product path against the reference's fp32 PyTorch path at <= 1e-5 instead of <= 1e-2 (BASELINE.json north_star).
// Not a performance path.
//
// tcgen05 has no fp32 MMA, so an fp32 GEMM runs on the SAME bf16 tcgen05 kernel through a 3-way split
//   x = x1 + x2 + x3,  w = w1 + w2 + w3      (bf16 pieces: 3 x 8 mantissa bits = fp32's 24)
//   x.w ~= x2w2 + x3w1 + x1w3 + x2w1 + x1w2 + x1w1      (dropped terms x2w3, x3w2, x3w3 <= 2^-24 |x||w|)
// laid out along the contraction dimension, SMALLEST products first so that they meet a small accumulator:
// A' = [x2|x3|x1|x2|x1|x1], B' = [w2|w1|w3|w1|w2|w1] (K' = 6 K), so ONE launch accumulates all six products in
// the fp32 TMEM accumulator (missm_expand6_bf16 builds A' / B').
// Attention, QuickGELU and the split run on the CUDA cores in fp32 with expf.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

// ---------------------------------------------------------------------------------------
// 3-way bf16 split of an fp32 matrix, six pieces in the order the GEMM operand needs
//   which = 0 (A operand): 2,3,1,2,1,1     which = 1 (B operand): 2,1,3,1,2,1
//   stack_rows = 0: dst[r, p * cols_pad + c]  (K-major operand, K = cols; columns cols..cols_pad-1 are zero)
//   stack_rows = 1: dst[p * rows + r, c]      (MN-major operand, K = rows)
// ---------------------------------------------------------------------------------------
__global__ void expand6_kernel(const float* __restrict__ src, long ld_src, int rows, int cols,
                               __nv_bfloat16* __restrict__ dst, long ld_dst, int cols_pad, int which, int stack_rows) {
  const long total = static_cast<long>(rows) * cols_pad;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols_pad), c = static_cast<int>(i % cols_pad);
    const float x = c < cols ? src[r * ld_src + c] : 0.f;
    __nv_bfloat16 pc[3];
    pc[0] = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(pc[0]);
    pc[1] = __float2bfloat16_rn(r1);
    pc[2] = __float2bfloat16_rn(r1 - __bfloat162float(pc[1]));
    const int pat_a[6] = {1, 2, 0, 1, 0, 0}, pat_b[6] = {1, 0, 2, 0, 1, 0};
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const __nv_bfloat16 v = pc[which == 0 ? pat_a[p] : pat_b[p]];
      if (stack_rows) dst[(static_cast<long>(p) * rows + r) * ld_dst + c] = v;
      else            dst[r * ld_dst + static_cast<long>(p) * cols_pad + c] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// QuickGELU in fp32 (transformers ACT2FN["quick_gelu"]: x * sigmoid(1.702 x))
// ---------------------------------------------------------------------------------------
__global__ void gelu_f32_fwd_kernel(const float* __restrict__ u, float* __restrict__ a, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float x = u[i];
    a[i] = x / (1.0f + expf(-1.702f * x));
  }
}
__global__ void gelu_f32_bwd_kernel(const float* __restrict__ d_a, const float* __restrict__ u, float* __restrict__ d_u, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float x = u[i];
    const float s = 1.0f / (1.0f + expf(-1.702f * x));
    d_u[i] = d_a[i] * (s + 1.702f * x * s * (1.0f - s));
  }
}

// ---------------------------------------------------------------------------------------
// fp32 attention, head_dim 64, one warp per query row (dQ, forward) or key row (dK, dV); any N <= 1024,
// causal / key-padding masks and strided sequences as missm_attention_fwd.  O(N^2 * 64) global loads per head.
// ---------------------------------------------------------------------------------------
constexpr int F32_MAXN = 1024;
constexpr int F32_WARPS = 4;

struct AttnF32Params {
  const float* qkv;
  long ld_qkv;
  int D, H, N, n_seq, s_in;
  long seq_outer, seq_inner, tok_stride;
  int causal;
  const int64_t* key_mask;
  const int32_t* mask_rows;
  int mask_div;
  float* out;
  long ld_o;
  float* lse;
  const float* d_out;
  float* delta;
  float* dqkv;
  float q_scale;
};

__device__ __forceinline__ long f32_row(const AttnF32Params& p, int s, int t) {
  return static_cast<long>(s / p.s_in) * p.seq_outer + static_cast<long>(s % p.s_in) * p.seq_inner +
         static_cast<long>(t) * p.tok_stride;
}
__device__ __forceinline__ const int64_t* f32_mask(const AttnF32Params& p, int s) {
  if (p.key_mask == nullptr) return nullptr;
  const int r = s / p.mask_div;
  return p.key_mask + static_cast<long>(p.mask_rows ? p.mask_rows[r] : r) * p.N;
}
// dot of a 64-float row in global memory with a 64-float vector in shared memory
__device__ __forceinline__ float dot64(const float* __restrict__ g, const float* __restrict__ sv) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float4 a = reinterpret_cast<const float4*>(g)[c];
    const float4 b = reinterpret_cast<const float4*>(sv)[c];
    s0 = fmaf(a.x, b.x, s0), s1 = fmaf(a.y, b.y, s1), s0 = fmaf(a.z, b.z, s0), s1 = fmaf(a.w, b.w, s1);
  }
  return s0 + s1;
}

// MODE 0: forward (out, lse).  MODE 1: dQ (+ delta).  Warp = query row.
template <int MODE>
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_f32_query_kernel(const AttnF32Params p) {
  __shared__ float sS[F32_WARPS][F32_MAXN];
  __shared__ __align__(16) float sQ[F32_WARPS][64];
  __shared__ __align__(16) float sG[F32_WARPS][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * F32_WARPS + warp, h = blockIdx.y, s = blockIdx.z;
  if (i >= p.N) return;                                   // warp-uniform
  const int64_t* km = f32_mask(p, s);
  const long qrow = f32_row(p, s, i);
  const float* q = p.qkv + qrow * p.ld_qkv + h * 64;
  sQ[warp][lane] = q[lane], sQ[warp][lane + 32] = q[lane + 32];
  float lse_i = 0.f, delta_i = 0.f;
  if (MODE == 1) {
    const float* g = p.d_out + qrow * p.ld_o + h * 64;
    const float* o = p.out + qrow * p.ld_o + h * 64;
    sG[warp][lane] = g[lane], sG[warp][lane + 32] = g[lane + 32];
    delta_i = warp_sum(g[lane] * o[lane] + g[lane + 32] * o[lane + 32]);
    lse_i = p.lse[(static_cast<long>(s) * p.H + h) * p.N + i];
    if (lane == 0) p.delta[(static_cast<long>(s) * p.H + h) * p.N + i] = delta_i;
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < p.N; j += 32) {
    const long krow = f32_row(p, s, j);
    const bool masked = (p.causal && j > i) || (km != nullptr && km[j] == 0);
    float v;
    if (MODE == 0) {
      v = masked ? -INFINITY : dot64(p.qkv + krow * p.ld_qkv + p.D + h * 64, sQ[warp]);
      mx = fmaxf(mx, v);
    } else {
      v = 0.f;
      if (!masked) {
        const float sc = dot64(p.qkv + krow * p.ld_qkv + p.D + h * 64, sQ[warp]);
        const float pr = expf(sc - lse_i);
        const float dp = dot64(p.qkv + krow * p.ld_qkv + 2 * p.D + h * 64, sG[warp]);
        v = pr * (dp - delta_i);                          // dS
      }
    }
    sS[warp][j] = v;
  }
  float sum = 0.f;
  if (MODE == 0) {
    mx = warp_max(mx);
    for (int j = lane; j < p.N; j += 32) {
      const float e = expf(sS[warp][j] - mx);             // exp(-inf) = 0 for masked keys
      sS[warp][j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
  }
  __syncwarp();
  // weighted sum over keys: lane owns columns 2*lane, 2*lane+1 of V (forward) or K (dQ)
  const int col = (MODE == 0 ? 2 * p.D : p.D) + h * 64 + 2 * lane;
  float a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < p.N; ++j) {
    const float w = sS[warp][j];
    const float2 x = *reinterpret_cast<const float2*>(p.qkv + f32_row(p, s, j) * p.ld_qkv + col);
    a0 = fmaf(w, x.x, a0), a1 = fmaf(w, x.y, a1);
  }
  if (MODE == 0) {
    const float inv = 1.0f / sum;
    *reinterpret_cast<float2*>(p.out + qrow * p.ld_o + h * 64 + 2 * lane) = make_float2(a0 * inv, a1 * inv);
    if (lane == 0 && p.lse != nullptr) p.lse[(static_cast<long>(s) * p.H + h) * p.N + i] = mx + logf(sum);
  } else {
    *reinterpret_cast<float2*>(p.dqkv + qrow * p.ld_qkv + h * 64 + 2 * lane) = make_float2(a0 * p.q_scale, a1 * p.q_scale);
  }
}

// dK, dV: warp = key row j; needs lse and delta of every query
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_f32_key_kernel(const AttnF32Params p) {
  __shared__ float sP[F32_WARPS][F32_MAXN];
  __shared__ float sD[F32_WARPS][F32_MAXN];
  __shared__ __align__(16) float sK[F32_WARPS][64];
  __shared__ __align__(16) float sV[F32_WARPS][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * F32_WARPS + warp, h = blockIdx.y, s = blockIdx.z;
  if (j >= p.N) return;                                   // warp-uniform
  const int64_t* km = f32_mask(p, s);
  const long krow = f32_row(p, s, j);
  const float* k = p.qkv + krow * p.ld_qkv + p.D + h * 64;
  const float* v = p.qkv + krow * p.ld_qkv + 2 * p.D + h * 64;
  sK[warp][lane] = k[lane], sK[warp][lane + 32] = k[lane + 32];
  sV[warp][lane] = v[lane], sV[warp][lane + 32] = v[lane + 32];
  __syncwarp();
  const bool key_masked = km != nullptr && km[j] == 0;
  const long sbase = (static_cast<long>(s) * p.H + h) * p.N;
  for (int i = lane; i < p.N; i += 32) {
    float pr = 0.f, ds = 0.f;
    if (!key_masked && !(p.causal && j > i)) {
      const long qrow = f32_row(p, s, i);
      const float sc = dot64(p.qkv + qrow * p.ld_qkv + h * 64, sK[warp]);
      pr = expf(sc - p.lse[sbase + i]);
      const float dp = dot64(p.d_out + qrow * p.ld_o + h * 64, sV[warp]);
      ds = pr * (dp - p.delta[sbase + i]);
    }
    sP[warp][i] = pr, sD[warp][i] = ds;
  }
  __syncwarp();
  float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
  for (int i = 0; i < p.N; ++i) {
    const long qrow = f32_row(p, s, i);
    const float2 g = *reinterpret_cast<const float2*>(p.d_out + qrow * p.ld_o + h * 64 + 2 * lane);
    const float2 q = *reinterpret_cast<const float2*>(p.qkv + qrow * p.ld_qkv + h * 64 + 2 * lane);
    const float pr = sP[warp][i], ds = sD[warp][i];
    v0 = fmaf(pr, g.x, v0), v1 = fmaf(pr, g.y, v1);
    k0 = fmaf(ds, q.x, k0), k1 = fmaf(ds, q.y, k1);
  }
  *reinterpret_cast<float2*>(p.dqkv + krow * p.ld_qkv + p.D + h * 64 + 2 * lane) = make_float2(k0, k1);
  *reinterpret_cast<float2*>(p.dqkv + krow * p.ld_qkv + 2 * p.D + h * 64 + 2 * lane) = make_float2(v0, v1);
}

static int fill_f32_params(const missm_attn_args* a, AttnF32Params& p) {
  MISSM_REQUIRE(a->head_dim == 64 && a->D == a->H * 64, "attention_f32: head_dim must be 64 (D=%d H=%d)", a->D, a->H);
  MISSM_REQUIRE(a->N >= 1 && a->N <= F32_MAXN, "attention_f32: N=%d out of range (<= %d)", a->N, F32_MAXN);
  MISSM_REQUIRE(a->ld_qkv % 4 == 0 && a->ld_o % 4 == 0, "attention_f32: leading dimensions must be multiples of 4");
  MISSM_REQUIRE(a->n_seq <= 65535 && a->H <= 65535, "attention_f32: n_seq=%d exceeds the grid (verification sizes only)", a->n_seq);
  p.qkv = static_cast<const float*>(a->qkv), p.ld_qkv = a->ld_qkv;
  p.D = a->D, p.H = a->H, p.N = a->N, p.n_seq = a->n_seq, p.s_in = a->s_in;
  p.seq_outer = a->seq_outer, p.seq_inner = a->seq_inner, p.tok_stride = a->tok_stride;
  p.causal = a->causal, p.key_mask = a->key_mask, p.mask_rows = a->mask_rows, p.mask_div = a->mask_div > 0 ? a->mask_div : 1;
  p.out = static_cast<float*>(a->out), p.ld_o = a->ld_o, p.lse = a->lse;
  p.d_out = static_cast<const float*>(a->d_out), p.delta = a->delta, p.dqkv = static_cast<float*>(a->dqkv);
  p.q_scale = a->q_scale;
  return 0;
}

}  // namespace missm

using namespace missm;

extern "C" int missm_expand6_bf16(const float* src, int64_t ld_src, int32_t rows, int32_t cols, void* dst,
                                  int64_t ld_dst, int32_t cols_pad, int32_t which, int32_t stack_rows, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  MISSM_REQUIRE(cols_pad >= cols && (which == 0 || which == 1), "expand6: cols_pad=%d cols=%d which=%d", cols_pad, cols, which);
  MISSM_REQUIRE(!stack_rows || cols_pad == cols, "expand6: row-stacked operands are not padded");
  const long total = static_cast<long>(rows) * cols_pad;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > 32 * kNumSMs) grid = 32 * kNumSMs;
  expand6_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ld_src, rows, cols, static_cast<__nv_bfloat16*>(dst),
                                                                       ld_dst, cols_pad, which, stack_rows); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_gelu_f32_fwd(const float* u, float* a, int64_t n, void* stream) {
  if (n == 0) return 0;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > 32 * kNumSMs) grid = 32 * kNumSMs;
  gelu_f32_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(u, a, n); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_gelu_f32_bwd(const float* d_a, const float* u, float* d_u, int64_t n, void* stream) {
  if (n == 0) return 0;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > 32 * kNumSMs) grid = 32 * kNumSMs;
  gelu_f32_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_a, u, d_u, n); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

/* qkv / out: fp32 (same layout contract as missm_attention_fwd) */
extern "C" int missm_attention_f32_fwd(const missm_attn_args* a, void* stream) {
  if (a->n_seq == 0) return 0;
  AttnF32Params p;
  if (int rc = fill_f32_params(a, p)) return rc;
  dim3 grid((p.N + F32_WARPS - 1) / F32_WARPS, p.H, p.n_seq);
  attn_f32_query_kernel<0><<<grid, F32_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(p); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

/* qkv / out / d_out / dqkv: fp32; lse from the forward; delta: workspace [n_seq, H, N] */
extern "C" int missm_attention_f32_bwd(const missm_attn_args* a, void* stream) {
  if (a->n_seq == 0) return 0;
  AttnF32Params p;
  if (int rc = fill_f32_params(a, p)) return rc;
  dim3 grid((p.N + F32_WARPS - 1) / F32_WARPS, p.H, p.n_seq);
  attn_f32_query_kernel<1><<<grid, F32_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(p); note_launch();
  attn_f32_key_kernel<<<grid, F32_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(p); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
