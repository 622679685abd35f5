// Fused multi-head attention, forward and backward, head_dim = 64, any sequence length,
// optional causal mask + key-padding mask.  Flash-style (scores never touch HBM).
//
// Replaces transformers 4.3x CLIPAttention's bmm -> (+causal) -> (+padding) -> softmax -> bmm
// chain (called at languagebind/image/modeling_image.py:121 [temporal, N = num_frames] and :140
// [spatial, N = 257 / 593; text N = 77 with the causal mask of :441-455 and the padding mask of
// :500-502]) and its autograd twin.  q arrives pre-scaled by head_dim^-0.5 (fused into the QKV
// GEMM epilogue), exactly like `q_proj(x) * self.scale`.
//
// Masking: the reference ADDS finfo.min for masked positions; every row keeps at least one
// unmasked key (the diagonal under the causal mask; key 0 = BOS is never padded), so the softmax
// weight of a masked key underflows to exactly 0 there.  This kernel sets those weights to
// exactly 0 as well.
//
// This is the general-shape path on the legacy tensor pipe (mma.sync m16n8k16 bf16, fp32
// accumulate, ldmatrix from padded shared memory, cp.async double buffering).  Sequences are
// addressed through (seq_outer, seq_inner, tok_stride) so the temporal attention of the video
// tower reads the [(b t) n d] activations in place -- the einops rearranges of
// modeling_image.py:112-118,127 are never materialised.
#include <cstdlib>

#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

constexpr int HD = 64;
constexpr int TILE = 64;          // rows per q tile and per kv tile
constexpr int SROW = HD + 8;      // padded smem row (144 B): conflict-free ldmatrix
constexpr int kAttnThreads = 128; // 4 warps x 16 rows
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  const __nv_bfloat16* qkv;  // [rows, ld_qkv]: q | k | v column blocks of width D
  long ld_qkv;
  int D, H, N;               // model width, heads, tokens per sequence
  int n_seq, s_in;           // sequences; addressing: base = (s / s_in)*seq_outer + (s % s_in)*seq_inner
  long seq_outer, seq_inner, tok_stride;
  int causal;
  const int64_t* key_mask;   // [*, N] (1 = attend) or null
  const int32_t* mask_rows;  // optional gather of key_mask rows (compaction), else identity
  int mask_div;              // key_mask row = (s / mask_div)
  __nv_bfloat16* out;        // fwd: [rows, ld_o]
  long ld_o;
  float* lse;                // [n_seq, H, N]
  // backward
  const __nv_bfloat16* d_out;  // [rows, ld_o]
  const float* delta;          // [n_seq, H, N]
  __nv_bfloat16* dqkv;         // [rows, ld_qkv]
  float q_scale;               // dq is multiplied by this (grad wrt the un-scaled q projection)
};

__device__ __forceinline__ long seq_base_row(const AttnParams& p, int s) {
  return static_cast<long>(s / p.s_in) * p.seq_outer + static_cast<long>(s % p.s_in) * p.seq_inner;
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// tile rows [t0, t0+64) of one head (64 columns at col_off) -> smem [64][SROW], zero-filled past N
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* stile, const __nv_bfloat16* base,
                                                long ld, long base_row, long tok_stride, int t0,
                                                int N, int col_off) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int chunk = threadIdx.x + i * kAttnThreads;
    const int r = chunk >> 3, c = (chunk & 7) * 8;
    const int t = t0 + r;
    const bool valid = t < N;
    const __nv_bfloat16* src = base + (base_row + static_cast<long>(valid ? t : 0) * tok_stride) * ld + col_off + c;
    cp_async_16(stile + r * SROW + c, src, valid);
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (16 rows of this warp x 64 k) from a [64][SROW] tile whose rows are the M index
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], const __nv_bfloat16* stile, int warp, int lane) {
  const int r = warp * 16 + (lane & 15), kc = (lane >> 4) * 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(f[kk], stile + r * SROW + kk * 16 + kc);
}

// acc[16 x 64] += A[16 x 64(k)] * T^T where the tile T is stored [n = 64][k = 64]  ("n-major rows")
// `npairs` (1..4, warp-uniform) = number of 16-row groups of the tile that hold valid rows: ragged
// sequence tails (N = 257 = 4*64 + 1) skip the all-padding groups instead of multiplying zeros.
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4],
                                            const __nv_bfloat16* stile, int lane, int npairs) {
  const int m = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      if (np >= npairs) continue;
      uint32_t b[4];
      ldsm_x4(b, stile + (np * 16 + (m >> 1) * 8 + r) * SROW + kk * 16 + (m & 1) * 8);
      mma_bf16(acc[2 * np], a[kk], b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a[kk], b[2], b[3]);
    }
  }
}
// acc[16 x 64(n)] += P[16 x 64(k)] * T where the tile T is stored [k = 64][n = 64]; P given as the
// fp32 accumulator fragments of a previous 16 x 64 product, converted to bf16 A fragments
__device__ __forceinline__ void mma_p_tile(float (&acc)[8][4], const float (&pacc)[8][4],
                                           const __nv_bfloat16* stile, int lane, int nksteps) {
  const int m = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (kk >= nksteps) continue;
    uint32_t a[4];
    a[0] = pack_bf16x2(pacc[2 * kk][0], pacc[2 * kk][1]);
    a[1] = pack_bf16x2(pacc[2 * kk][2], pacc[2 * kk][3]);
    a[2] = pack_bf16x2(pacc[2 * kk + 1][0], pacc[2 * kk + 1][1]);
    a[3] = pack_bf16x2(pacc[2 * kk + 1][2], pacc[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, stile + (kk * 16 + (m & 1) * 8 + r) * SROW + np * 16 + (m >> 1) * 8);
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

__device__ __forceinline__ const int64_t* mask_row_ptr(const AttnParams& p, int s) {
  if (p.key_mask == nullptr) return nullptr;
  const int b = s / p.mask_div;
  const long row = p.mask_rows ? p.mask_rows[b] : b;
  return p.key_mask + row * p.N;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const AttnParams p) {
  __shared__ __align__(16) __nv_bfloat16 sQ[TILE * SROW];
  __shared__ __align__(16) __nv_bfloat16 sK[2][TILE * SROW];
  __shared__ __align__(16) __nv_bfloat16 sV[2][TILE * SROW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, s = blockIdx.z;
  const long base = seq_base_row(p, s);
  const int64_t* kmask = mask_row_ptr(p, s);

  load_tile_async(sQ, p.qkv, p.ld_qkv, base, p.tok_stride, q0, p.N, h * HD);
  cp_async_commit();
  int nkv = (p.N + TILE - 1) / TILE;
  if (p.causal) nkv = min(nkv, (q0 + TILE - 1) / TILE + 1);
  load_tile_async(sK[0], p.qkv, p.ld_qkv, base, p.tok_stride, 0, p.N, p.D + h * HD);
  load_tile_async(sV[0], p.qkv, p.ld_qkv, base, p.tok_stride, 0, p.N, 2 * p.D + h * HD);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[4][4];
  load_a_frags(qf, sQ, warp, lane);

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  const bool warp_active = (q0 + warp * 16) < p.N;   // strips past the sequence end only help loading

  for (int j = 0; j < nkv; ++j) {
    if (j + 1 < nkv) {
      load_tile_async(sK[(j + 1) & 1], p.qkv, p.ld_qkv, base, p.tok_stride, (j + 1) * TILE, p.N, p.D + h * HD);
      load_tile_async(sV[(j + 1) & 1], p.qkv, p.ld_qkv, base, p.tok_stride, (j + 1) * TILE, p.N, 2 * p.D + h * HD);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (warp_active) {
    const int ngrp = min(4, (p.N - j * TILE + 15) / 16);   // valid 16-row groups of this kv tile
    float sc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
    mma_a_tileT(sc, qf, sK[j & 1], lane, ngrp);

    // masking + online softmax (base-2 domain)
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kv = j * TILE + n * 8 + t4 * 2 + (e & 1);
        const int qr = qrow[e >> 1];
        bool ok = kv < p.N;
        if (p.causal) ok = ok && (kv <= qr);
        if (kmask != nullptr && kv < p.N) ok = ok && (kmask[kv] != 0);
        const float v = ok ? sc[n][e] * kLog2e : -INFINITY;
        sc[n][e] = v;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
      }
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float m_new = fmaxf(m_run[r], tmax[r]);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = exp2f(m_run[r] - m_use[r]);  // m_run = -inf -> 0
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(sc[n][e] - m_use[e >> 1]);
        sc[n][e] = pv;
        psum[e >> 1] += pv;
      }
      o[n][0] *= corr[0], o[n][1] *= corr[0], o[n][2] *= corr[1], o[n][3] *= corr[1];
    }
    l_run[0] += psum[0], l_run[1] += psum[1];
    mma_p_tile(o, sc, sV[j & 1], lane, ngrp);
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qr = qrow[r];
    if (qr >= p.N) continue;
    const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
    __nv_bfloat16* orow = p.out + (base + static_cast<long>(qr) * p.tok_stride) * p.ld_o + h * HD;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      *reinterpret_cast<uint32_t*>(orow + n * 8 + t4 * 2) =
          pack_bf16x2(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
    if (t4 == 0 && p.lse != nullptr)
      p.lse[(static_cast<long>(s) * p.H + h) * p.N + qr] = (m_run[r] + log2f(l_run[r])) / kLog2e;
  }
}

// ---------------------------------------------------------------------------------------
// backward preprocess: delta[s, h, t] = sum_d dO * O
// ---------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const AttnParams p) {
  const long total = static_cast<long>(p.n_seq) * p.H * p.N;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % p.N);
    const int h = static_cast<int>((i / p.N) % p.H);
    const int s = static_cast<int>(i / (static_cast<long>(p.N) * p.H));
    const long row = seq_base_row(p, s) + static_cast<long>(t) * p.tok_stride;
    const uint4* a = reinterpret_cast<const uint4*>(p.out + row * p.ld_o + h * HD);
    const uint4* b = reinterpret_cast<const uint4*>(p.d_out + row * p.ld_o + h * HD);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 x = a[c], y = b[c];
      float2 x0 = unpack_bf16x2(x.x), x1 = unpack_bf16x2(x.y), x2 = unpack_bf16x2(x.z), x3 = unpack_bf16x2(x.w);
      float2 y0 = unpack_bf16x2(y.x), y1 = unpack_bf16x2(y.y), y2 = unpack_bf16x2(y.z), y3 = unpack_bf16x2(y.w);
      acc += x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y + x2.x * y2.x + x2.y * y2.y + x3.x * y3.x +
             x3.y * y3.y;
    }
    const_cast<float*>(p.delta)[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------
// backward, dQ: one CTA per (q tile, head, sequence); loops over kv tiles
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dq_kernel(const AttnParams p) {
  __shared__ __align__(16) __nv_bfloat16 sK[2][TILE * SROW];
  __shared__ __align__(16) __nv_bfloat16 sV[2][TILE * SROW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, s = blockIdx.z;
  const long base = seq_base_row(p, s);
  const int64_t* kmask = mask_row_ptr(p, s);

  // stage Q and dO through the (not yet used) second K/V buffers
  load_tile_async(sK[1], p.qkv, p.ld_qkv, base, p.tok_stride, q0, p.N, h * HD);
  load_tile_async(sV[1], p.d_out, p.ld_o, base, p.tok_stride, q0, p.N, h * HD);
  cp_async_commit();
  int nkv = (p.N + TILE - 1) / TILE;
  if (p.causal) nkv = min(nkv, (q0 + TILE - 1) / TILE + 1);
  load_tile_async(sK[0], p.qkv, p.ld_qkv, base, p.tok_stride, 0, p.N, p.D + h * HD);
  load_tile_async(sV[0], p.qkv, p.ld_qkv, base, p.tok_stride, 0, p.N, 2 * p.D + h * HD);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[4][4], dof[4][4];
  load_a_frags(qf, sK[1], warp, lane);
  load_a_frags(dof, sV[1], warp, lane);
  __syncthreads();  // everyone has its fragments before buffer 1 is refilled

  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  float lse2[2], dlt[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const bool ok = qrow[r] < p.N;
    const long li = (static_cast<long>(s) * p.H + h) * p.N + (ok ? qrow[r] : 0);
    lse2[r] = ok ? p.lse[li] * kLog2e : INFINITY;
    dlt[r] = ok ? p.delta[li] : 0.f;
  }
  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
  const bool warp_active = (q0 + warp * 16) < p.N;

  for (int j = 0; j < nkv; ++j) {
    if (j + 1 < nkv) {
      load_tile_async(sK[(j + 1) & 1], p.qkv, p.ld_qkv, base, p.tok_stride, (j + 1) * TILE, p.N, p.D + h * HD);
      load_tile_async(sV[(j + 1) & 1], p.qkv, p.ld_qkv, base, p.tok_stride, (j + 1) * TILE, p.N, 2 * p.D + h * HD);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (warp_active) {
    const int ngrp = min(4, (p.N - j * TILE + 15) / 16);
    float sc[8][4], dp[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
    mma_a_tileT(sc, qf, sK[j & 1], lane, ngrp);   // S  = Q K^T
    mma_a_tileT(dp, dof, sV[j & 1], lane, ngrp);  // dP = dO V^T
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kv = j * TILE + n * 8 + t4 * 2 + (e & 1);
        const int r = e >> 1;
        bool ok = kv < p.N;
        if (p.causal) ok = ok && (kv <= qrow[r]);
        if (kmask != nullptr && kv < p.N) ok = ok && (kmask[kv] != 0);
        const float pv = ok ? exp2f(sc[n][e] * kLog2e - lse2[r]) : 0.f;
        sc[n][e] = pv * (dp[n][e] - dlt[r]);  // dS
      }
    }
    mma_p_tile(dq, sc, sK[j & 1], lane, ngrp);  // dQ += dS K
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (qrow[r] >= p.N) continue;
    __nv_bfloat16* drow = p.dqkv + (base + static_cast<long>(qrow[r]) * p.tok_stride) * p.ld_qkv + h * HD;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      *reinterpret_cast<uint32_t*>(drow + n * 8 + t4 * 2) =
          pack_bf16x2(dq[n][2 * r] * p.q_scale, dq[n][2 * r + 1] * p.q_scale);
  }
}

// ---------------------------------------------------------------------------------------
// backward, dK / dV: one CTA per (kv tile, head, sequence); loops over q tiles, works on the
// transposed scores S^T = K Q^T so that kv is the accumulator row index
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dkv_kernel(const AttnParams p) {
  __shared__ __align__(16) __nv_bfloat16 sQ[2][TILE * SROW];
  __shared__ __align__(16) __nv_bfloat16 sdO[2][TILE * SROW];
  __shared__ float sLse[2][TILE], sDelta[2][TILE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int k0 = blockIdx.x * TILE, h = blockIdx.y, s = blockIdx.z;
  const long base = seq_base_row(p, s);
  const int64_t* kmask = mask_row_ptr(p, s);
  const long stat_base = (static_cast<long>(s) * p.H + h) * p.N;

  load_tile_async(sQ[1], p.qkv, p.ld_qkv, base, p.tok_stride, k0, p.N, p.D + h * HD);       // K
  load_tile_async(sdO[1], p.qkv, p.ld_qkv, base, p.tok_stride, k0, p.N, 2 * p.D + h * HD);  // V
  cp_async_commit();
  const int nq = (p.N + TILE - 1) / TILE;
  const int jq0 = p.causal ? k0 / TILE : 0;  // q tiles entirely above the diagonal contribute 0
  auto load_q_tile = [&](int j, int buf) {
    load_tile_async(sQ[buf], p.qkv, p.ld_qkv, base, p.tok_stride, j * TILE, p.N, h * HD);
    load_tile_async(sdO[buf], p.d_out, p.ld_o, base, p.tok_stride, j * TILE, p.N, h * HD);
    if (threadIdx.x < TILE) {
      const int t = j * TILE + threadIdx.x;
      sLse[buf][threadIdx.x] = t < p.N ? p.lse[stat_base + t] * kLog2e : INFINITY;
      sDelta[buf][threadIdx.x] = t < p.N ? p.delta[stat_base + t] : 0.f;
    }
  };
  load_q_tile(jq0, 0);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(kf, sQ[1], warp, lane);
  load_a_frags(vf, sdO[1], warp, lane);
  __syncthreads();

  const int kvrow[2] = {k0 + warp * 16 + g, k0 + warp * 16 + g + 8};
  bool kv_ok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    kv_ok[r] = kvrow[r] < p.N;
    if (kmask != nullptr && kv_ok[r]) kv_ok[r] = kmask[kvrow[r]] != 0;
  }
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  const bool warp_active = (k0 + warp * 16) < p.N;

  for (int j = jq0; j < nq; ++j) {
    const int buf = (j - jq0) & 1;
    if (j + 1 < nq) {
      load_q_tile(j + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (warp_active) {
    const int ngrp = min(4, (p.N - j * TILE + 15) / 16);   // valid 16-row groups of this q tile
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
      dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
    }
    mma_a_tileT(st, kf, sQ[buf], lane, ngrp);    // S^T  = K Q^T
    mma_a_tileT(dpt, vf, sdO[buf], lane, ngrp);  // dP^T = V dO^T
    float pt[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = n * 8 + t4 * 2 + (e & 1);
        const int qt = j * TILE + qi;
        const int r = e >> 1;
        bool ok = kv_ok[r];
        if (p.causal) ok = ok && (kvrow[r] <= qt);
        const float pv = ok ? exp2f(st[n][e] * kLog2e - sLse[buf][qi]) : 0.f;  // lse=+inf past N -> 0
        pt[n][e] = pv;
        st[n][e] = pv * (dpt[n][e] - sDelta[buf][qi]);  // dS^T
      }
    }
    mma_p_tile(dv, pt, sdO[buf], lane, ngrp);  // dV += P^T dO
    mma_p_tile(dk, st, sQ[buf], lane, ngrp);   // dK += dS^T Q
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (kvrow[r] >= p.N) continue;
    __nv_bfloat16* drow = p.dqkv + (base + static_cast<long>(kvrow[r]) * p.tok_stride) * p.ld_qkv;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<uint32_t*>(drow + p.D + h * HD + n * 8 + t4 * 2) = pack_bf16x2(dk[n][2 * r], dk[n][2 * r + 1]);
      *reinterpret_cast<uint32_t*>(drow + 2 * p.D + h * HD + n * 8 + t4 * 2) = pack_bf16x2(dv[n][2 * r], dv[n][2 * r + 1]);
    }
  }
}

static int fill_params(AttnParams& p, const missm_attn_args* a) {
  MISSM_REQUIRE(a->head_dim == HD, "attention: head_dim=%d (only 64 is built)", a->head_dim);
  MISSM_REQUIRE(a->D == a->H * HD, "attention: D=%d != H*64 (H=%d)", a->D, a->H);
  MISSM_REQUIRE(a->ld_qkv % 8 == 0 && a->ld_o % 8 == 0, "attention: leading dims must be multiples of 8");
  MISSM_REQUIRE(a->s_in >= 1 && a->N >= 1, "attention: bad sequence description");
  p.qkv = static_cast<const __nv_bfloat16*>(a->qkv);
  p.ld_qkv = a->ld_qkv;
  p.D = a->D, p.H = a->H, p.N = a->N;
  p.n_seq = a->n_seq, p.s_in = a->s_in;
  p.seq_outer = a->seq_outer, p.seq_inner = a->seq_inner, p.tok_stride = a->tok_stride;
  p.causal = a->causal;
  p.key_mask = a->key_mask, p.mask_rows = a->mask_rows, p.mask_div = a->mask_div > 0 ? a->mask_div : 1;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.ld_o = a->ld_o;
  p.lse = a->lse;
  p.d_out = static_cast<const __nv_bfloat16*>(a->d_out);
  p.delta = a->delta;
  p.dqkv = static_cast<__nv_bfloat16*>(a->dqkv);
  p.q_scale = a->q_scale;
  return 0;
}

}  // namespace missm

using namespace missm;

namespace missm {
int attention_fwd_tc(const missm_attn_args* a, cudaStream_t stream);  // attention_tc.cu
int attention_bwd_tc(const missm_attn_args* a, cudaStream_t stream);  // attention_tc_bwd.cu
int attention_fwd_small(const missm_attn_args* a, cudaStream_t stream);   // attention_small.cu (N <= 8)
int attention_bwd_small(const missm_attn_args* a, cudaStream_t stream);
}

extern "C" int missm_attention_fwd(const missm_attn_args* a, void* stream) {
  if (a->n_seq == 0) return 0;
  MISSM_REQUIRE(a->qkv && a->out, "attention_fwd: null tensor");
  {   // short sequences (temporal attention of the video tower): one warp per (sequence, head)
    const int rc = attention_fwd_small(a, static_cast<cudaStream_t>(stream));
    if (rc >= 0) return rc;
  }
  // tcgen05 path for the shapes it covers (spatial ViT attention); general shapes below
  static const bool legacy_only = getenv("MISSM_ATTN_LEGACY") != nullptr || getenv("MISSM_ATTN_LEGACY_FWD") != nullptr;
  if (!legacy_only) {
    MISSM_REQUIRE(a->qkv && a->out, "attention_fwd: null tensor");
    const int rc = attention_fwd_tc(a, static_cast<cudaStream_t>(stream));
    if (rc >= 0) return rc;
  }
  AttnParams p;
  if (int rc = fill_params(p, a)) return rc;
  MISSM_REQUIRE(p.qkv && p.out, "attention_fwd: null tensor");
  dim3 grid((p.N + TILE - 1) / TILE, p.H, p.n_seq);
  attn_fwd_kernel<<<grid, kAttnThreads, 0, static_cast<cudaStream_t>(stream)>>>(p); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_attention_bwd(missm_attn_args* a, void* stream) {
  a->colsum_done = 0;
  if (a->n_seq == 0) return 0;
  AttnParams p;
  if (int rc = fill_params(p, a)) return rc;
  MISSM_REQUIRE(p.qkv && p.out && p.d_out && p.lse && p.delta && p.dqkv, "attention_bwd: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long total = static_cast<long>(p.n_seq) * p.H * p.N;
  int dgrid = static_cast<int>((total + 255) / 256);
  if (dgrid > 16 * kNumSMs) dgrid = 16 * kNumSMs;
  {
    const int rc = attention_bwd_small(a, st);    // short sequences: one warp per (sequence, head), no delta pass
    if (rc >= 0) return rc;
  }
  static const bool legacy_only = getenv("MISSM_ATTN_LEGACY") != nullptr || getenv("MISSM_ATTN_LEGACY_BWD") != nullptr;
  if (!legacy_only) {
    const int rc = attention_bwd_tc(a, st);   // tcgen05 path for the shapes it covers (computes delta itself)
    if (rc >= 0) return rc;
  }
  attn_delta_kernel<<<dgrid, 256, 0, st>>>(p); note_launch();
  dim3 grid((p.N + TILE - 1) / TILE, p.H, p.n_seq);
  attn_bwd_dkv_kernel<<<grid, kAttnThreads, 0, st>>>(p); note_launch();
  attn_bwd_dq_kernel<<<grid, kAttnThreads, 0, st>>>(p); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
