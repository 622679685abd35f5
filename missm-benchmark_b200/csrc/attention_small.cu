// Attention over SHORT sequences (N <= 8 tokens, head_dim 64, no mask): the temporal attention of the video tower
// (languagebind/video/modeling_video.py:211-233 = CLIPEncoderLayer.forward modeling_image.py:105-127: sequences of
// T = 8 frames, one per (sample, patch) -- 7 400 sequences x 16 heads per layer at B = 32), forward and backward.
//
// The op is HBM-bound: per (sequence, head) it reads 3 x 8 rows of 128 bytes and does 4 * 8 * 8 * 64 = 16 K flops
// -- nothing for a tensor core to win, and the generic mma.sync kernel spent a 64-row tile and a whole CTA on
// every 8-row sequence (forward 746 us, backward 2.6 ms per layer at the configs[2] shape; ncu launch list
// profiles/r02g_launch_summary_config2_1layer.txt).  Here ONE WARP owns a (sequence, head):
//   * lane l holds dimensions 2l, 2l+1 of every q / k / v (/ dO) row: 4-byte loads, a warp reads one contiguous
//     128-byte segment per row; adjacent warps take adjacent heads of the same rows;
//   * the 64 scores are 64 two-term partial products per lane, summed across lanes with a butterfly
//     transpose-reduce (31 shuffles per 32 values) that leaves score (i, j) in lane 8 (i % 4) + j, slot i / 4;
//   * softmax over j = 3 xor-shuffles inside groups of 8 lanes; probabilities are broadcast back with shuffles for
//     the 8 x 8 x 2 accumulation FMAs of every lane.
// The backward recomputes S and P (no lse needed), dP = dO V^T the same way, dS = P o (dP - rowsum(P o dP)), and
// writes dq (x q_scale), dk, dv.  Sequence addressing: row(s, t) = (s / s_in) * seq_outer + (s % s_in) * seq_inner +
// t * tok_stride, as everywhere in this library -- the `(b t) n d <-> (b n) t d` rearranges of the reference
// (modeling_image.py:112-118,127) are never materialised.
// Algorithmic bytes per (sequence, head): forward 3 * N * 128 read + N * 128 written; backward 4 * N * 128 read +
// 3 * N * 128 written.
#include <cstdlib>

#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

constexpr int SA_T = 8;          // tokens per sequence (max)
constexpr int SA_WARPS = 8;      // warps per CTA
constexpr float kLog2eSa = 1.4426950408889634f;

struct SmallAttnParams {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  float* lse;
  const __nv_bfloat16* d_out;
  __nv_bfloat16* dqkv;
  long ld_qkv, ld_o;
  long seq_outer, seq_inner, tok_stride;
  int D, H, N, n_seq, s_in;
  float q_scale;
};

__device__ __forceinline__ float2 ld_bf2(const __nv_bfloat16* p) {
  return unpack_bf16x2(__ldg(reinterpret_cast<const unsigned int*>(p)));
}

// v[0..63]: this lane's partial of the 64 products (index p = i * 8 + j).  Returns the two totals this lane owns:
// .x = total of p = lane (i = lane / 8, j = lane % 8), .y = total of p = lane + 32 (i = 4 + lane / 8).
__device__ __forceinline__ float2 reduce64(float (&v)[64], int lane) {
  float lo[32], hi[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) lo[i] = v[i], hi[i] = v[i + 32];
  float2 r;
  r.x = warp_colsum32(lo, lane);
  r.y = warp_colsum32(hi, lane);
  return r;
}

// partial products of this lane: v[i * 8 + j] = a[i] . b[j] over the lane's two dimensions
__device__ __forceinline__ void outer8(const float2 (&a)[SA_T], const float2 (&b)[SA_T], float (&v)[64]) {
#pragma unroll
  for (int i = 0; i < SA_T; ++i)
#pragma unroll
    for (int j = 0; j < SA_T; ++j) v[i * 8 + j] = fmaf(a[i].x, b[j].x, a[i].y * b[j].y);
}

// softmax over j (the 8 lanes of a group) of the two score rows this lane takes part in; columns j >= N masked
__device__ __forceinline__ float2 softmax8(float2 s, int lane, int N, float2* lse) {
  const bool ok = (lane & 7) < N;
  float mx0 = ok ? s.x : -INFINITY, mx1 = ok ? s.y : -INFINITY;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
  }
  float e0 = ok ? fast_ex2((s.x - mx0) * kLog2eSa) : 0.f, e1 = ok ? fast_ex2((s.y - mx1) * kLog2eSa) : 0.f;
  float z0 = e0, z1 = e1;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    z0 += __shfl_xor_sync(0xffffffffu, z0, o);
    z1 += __shfl_xor_sync(0xffffffffu, z1, o);
  }
  if (lse != nullptr) *lse = make_float2(mx0 + __logf(z0), mx1 + __logf(z1));
  return make_float2(e0 / z0, e1 / z1);
}

// acc[i] += sum_j w(i, j) * b[j]   with w(i, j) held by lane 8 (i % 4) + j in slot i / 4 (.x: i < 4, .y: i >= 4)
__device__ __forceinline__ void apply_rows(float2 w, const float2 (&b)[SA_T], float2 (&acc)[SA_T]) {
#pragma unroll
  for (int i = 0; i < SA_T; ++i)
#pragma unroll
    for (int j = 0; j < SA_T; ++j) {
      const float x = __shfl_sync(0xffffffffu, i < 4 ? w.x : w.y, 8 * (i & 3) + j);
      acc[i].x = fmaf(x, b[j].x, acc[i].x), acc[i].y = fmaf(x, b[j].y, acc[i].y);
    }
}
// acc[j] += sum_i w(i, j) * a[i]   (the transposed application)
__device__ __forceinline__ void apply_cols(float2 w, const float2 (&a)[SA_T], float2 (&acc)[SA_T]) {
#pragma unroll
  for (int i = 0; i < SA_T; ++i)
#pragma unroll
    for (int j = 0; j < SA_T; ++j) {
      const float x = __shfl_sync(0xffffffffu, i < 4 ? w.x : w.y, 8 * (i & 3) + j);
      acc[j].x = fmaf(x, a[i].x, acc[j].x), acc[j].y = fmaf(x, a[i].y, acc[j].y);
    }
}

template <bool BWD>
__global__ void __launch_bounds__(SA_WARPS * 32)
attn_small_kernel(const SmallAttnParams p) {
  const int lane = threadIdx.x & 31;
  const long w = static_cast<long>(blockIdx.x) * SA_WARPS + (threadIdx.x >> 5);
  if (w >= static_cast<long>(p.n_seq) * p.H) return;     // warp-uniform
  const int s = static_cast<int>(w / p.H), h = static_cast<int>(w % p.H);
  const long base = static_cast<long>(s / p.s_in) * p.seq_outer + static_cast<long>(s % p.s_in) * p.seq_inner;
  const int col = h * 64 + 2 * lane;

  float2 q[SA_T], k[SA_T], v[SA_T];
#pragma unroll
  for (int t = 0; t < SA_T; ++t) {
    q[t] = k[t] = v[t] = make_float2(0.f, 0.f);
    if (t < p.N) {
      const __nv_bfloat16* r = p.qkv + (base + t * p.tok_stride) * p.ld_qkv + col;
      q[t] = ld_bf2(r), k[t] = ld_bf2(r + p.D), v[t] = ld_bf2(r + 2 * p.D);
    }
  }
  float part[64];
  outer8(q, k, part);
  const float2 sc = reduce64(part, lane);          // q is pre-scaled by head_dim^-0.5 (GEMM epilogue)
  float2 lse;
  const float2 pr = softmax8(sc, lane, p.N, &lse);

  if constexpr (!BWD) {
    float2 o[SA_T];
#pragma unroll
    for (int t = 0; t < SA_T; ++t) o[t] = make_float2(0.f, 0.f);
    apply_rows(pr, v, o);
#pragma unroll
    for (int t = 0; t < SA_T; ++t)
      if (t < p.N)
        *reinterpret_cast<unsigned int*>(p.out + (base + t * p.tok_stride) * p.ld_o + col) = pack_bf16x2(o[t].x, o[t].y);
    if (p.lse != nullptr && (lane & 7) == 0) {
      const int i0 = lane >> 3;
      float* l = p.lse + w * p.N;
      if (i0 < p.N) l[i0] = lse.x;
      if (i0 + 4 < p.N) l[i0 + 4] = lse.y;
    }
  } else {
    float2 d_o[SA_T];
#pragma unroll
    for (int t = 0; t < SA_T; ++t) {
      d_o[t] = make_float2(0.f, 0.f);
      if (t < p.N) d_o[t] = ld_bf2(p.d_out + (base + t * p.tok_stride) * p.ld_o + col);
    }
    outer8(d_o, v, part);
    const float2 dp = reduce64(part, lane);        // dP(i, j) = dO_i . v_j
    float de0 = pr.x * dp.x, de1 = pr.y * dp.y;    // delta_i = sum_j P dP
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      de0 += __shfl_xor_sync(0xffffffffu, de0, o);
      de1 += __shfl_xor_sync(0xffffffffu, de1, o);
    }
    const float2 ds = make_float2(pr.x * (dp.x - de0), pr.y * (dp.y - de1));
    float2 dq[SA_T], dk[SA_T], dv[SA_T];
#pragma unroll
    for (int t = 0; t < SA_T; ++t) dq[t] = dk[t] = dv[t] = make_float2(0.f, 0.f);
    apply_rows(ds, k, dq);          // dq_i = sum_j dS_ij k_j        (x q_scale below: gradient w.r.t. the un-scaled q)
    apply_cols(ds, q, dk);          // dk_j = sum_i dS_ij q_i        (q as stored, i.e. scaled)
    apply_cols(pr, d_o, dv);        // dv_j = sum_i P_ij dO_i
#pragma unroll
    for (int t = 0; t < SA_T; ++t)
      if (t < p.N) {
        __nv_bfloat16* r = p.dqkv + (base + t * p.tok_stride) * p.ld_qkv + col;
        *reinterpret_cast<unsigned int*>(r) = pack_bf16x2(dq[t].x * p.q_scale, dq[t].y * p.q_scale);
        *reinterpret_cast<unsigned int*>(r + p.D) = pack_bf16x2(dk[t].x, dk[t].y);
        *reinterpret_cast<unsigned int*>(r + 2 * p.D) = pack_bf16x2(dv[t].x, dv[t].y);
      }
  }
}

static bool small_ok(const missm_attn_args* a) {
  static const bool off = getenv("MISSM_ATTN_NO_SMALL") != nullptr;     // A/B switch: back to the generic kernel
  return !off && a->N <= SA_T && !a->causal && a->key_mask == nullptr && a->head_dim == 64 && a->D == a->H * 64;
}
static void fill(SmallAttnParams& p, const missm_attn_args* a) {
  p.qkv = static_cast<const __nv_bfloat16*>(a->qkv), p.out = static_cast<__nv_bfloat16*>(a->out), p.lse = a->lse;
  p.d_out = static_cast<const __nv_bfloat16*>(a->d_out), p.dqkv = static_cast<__nv_bfloat16*>(a->dqkv);
  p.ld_qkv = a->ld_qkv, p.ld_o = a->ld_o;
  p.seq_outer = a->seq_outer, p.seq_inner = a->seq_inner, p.tok_stride = a->tok_stride;
  p.D = a->D, p.H = a->H, p.N = a->N, p.n_seq = a->n_seq, p.s_in = a->s_in, p.q_scale = a->q_scale;
}

// return 0 if launched, -1 if the shape is not handled here, > 0 on error
int attention_fwd_small(const missm_attn_args* a, cudaStream_t stream) {
  if (!small_ok(a)) return -1;
  SmallAttnParams p;
  fill(p, a);
  const long warps = static_cast<long>(a->n_seq) * a->H;
  attn_small_kernel<false><<<static_cast<unsigned>((warps + SA_WARPS - 1) / SA_WARPS), SA_WARPS * 32, 0, stream>>>(p); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int attention_bwd_small(const missm_attn_args* a, cudaStream_t stream) {
  if (!small_ok(a)) return -1;
  SmallAttnParams p;
  fill(p, a);
  const long warps = static_cast<long>(a->n_seq) * a->H;
  attn_small_kernel<true><<<static_cast<unsigned>((warps + SA_WARPS - 1) / SA_WARPS), SA_WARPS * 32, 0, stream>>>(p); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace missm
