// The odd token of the ViT sequences (N = 257 = 2 * 128 + 1) on the CUDA cores.
//
// The tcgen05 attention kernels walk 128-row tiles; with N = 257 a third tile would carry ONE valid
// row and cost as much as a full one (the softmax warps own one row per thread, so their time does not
// shrink with the number of valid rows).  Instead, when N % 128 == 1 the tiles cover rows [0, N-1) and a
// team of two otherwise idle warps of the CTA computes everything that belongs to row N-1 with plain FMAs, straight from the
// 128B-swizzled shared-memory operands TMA already brought in for the tensor path:
//   forward : out[N-1] = softmax(q K^T) V                       (K, V resident)
//   dQ      : dQ[N-1]  = (P o (dO V^T - delta)) K               (K, V resident)
//   dKV     : dK[N-1]  = dS[:, N-1]^T Q,  dV[N-1] = P[:, N-1]^T dO   (Q, dO resident)
// About 3-6 k instructions per item, split over the two warps and overlapped with the tensor pipeline's two
// tiles (a single warp is too slow: latency-bound at ~0.25 IPC it became the critical path of the item).
#pragma once
#include <cstdlib>

#include "missm_common.cuh"

namespace missm {

// host: N = k * 128 + 1 (MISSM_ATTN_NO_TAIL=1 restores the third tile, for A/B measurements)
inline bool attention_tail_enabled(int N) {
  static const bool off = getenv("MISSM_ATTN_NO_TAIL") != nullptr;
  return !off && N > 128 && N % 128 == 1;
}

// resident operand: rows of 64 bf16 = 128 B; TMA's 128B swizzle stores the 16-byte chunk c of row r at
// chunk position c ^ (r & 7) (buffers are 1024-byte aligned)
__device__ __forceinline__ uint32_t tail_chunk_addr(uint32_t base, int r, int c) {
  return base + r * 128 + ((c ^ (r & 7)) << 4);
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

// a 64-element bf16 row in global memory -> fp32 registers (all lanes load the same 128 bytes)
__device__ __forceinline__ void tail_load_row64(const __nv_bfloat16* g, float (&a)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(g) + c);
    const float2 x = unpack_bf16x2(q.x), y = unpack_bf16x2(q.y), z = unpack_bf16x2(q.z), w = unpack_bf16x2(q.w);
    a[8 * c] = x.x, a[8 * c + 1] = x.y, a[8 * c + 2] = y.x, a[8 * c + 3] = y.y;
    a[8 * c + 4] = z.x, a[8 * c + 5] = z.y, a[8 * c + 6] = w.x, a[8 * c + 7] = w.y;
  }
}

// a . R[r]   (one thread, one row; 8 conflict-free 16-byte reads: threads r..r+7 cover all 8 chunk
// positions).  Four independent accumulation chains.
__device__ __forceinline__ float tail_dot64(uint32_t base, int r, const float (&a)[64]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const uint32_t rbase = base + r * 128;
  const int x7 = r & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 q = lds128(rbase + ((c ^ x7) << 4));
    const float2 x = unpack_bf16x2(q.x), y = unpack_bf16x2(q.y), z = unpack_bf16x2(q.z), w = unpack_bf16x2(q.w);
    s0 = fmaf(a[8 * c], x.x, s0), s1 = fmaf(a[8 * c + 1], x.y, s1);
    s2 = fmaf(a[8 * c + 2], y.x, s2), s3 = fmaf(a[8 * c + 3], y.y, s3);
    s0 = fmaf(a[8 * c + 4], z.x, s0), s1 = fmaf(a[8 * c + 5], z.y, s1);
    s2 = fmaf(a[8 * c + 6], w.x, s2), s3 = fmaf(a[8 * c + 7], w.y, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// acc[j] = sum_{r0 <= r < r1} w[r] * R[r][8 * (lane & 7) + j]   (valid in lanes 0..7 = the 8 chunks of a row)
// lane = (row group rg = lane >> 3, chunk = lane & 7): each step reads 4 whole rows (conflict-free), two
// steps in flight;  w: fp32 weights in shared memory (zero for rows that must not count).
__device__ __forceinline__ void tail_weighted_rowsum(uint32_t base, uint32_t w_saddr, int r0, int r1, int lane,
                                                     float (&acc)[8]) {
  const int ch = lane & 7, rg = lane >> 3;
  float b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = b[j] = 0.f;
  int r = r0 + rg;
  for (; r + 4 < r1; r += 8) {
    const float w0 = lds_f32(w_saddr + r * 4), w1 = lds_f32(w_saddr + (r + 4) * 4);
    const uint4 q0 = lds128(tail_chunk_addr(base, r, ch)), q1 = lds128(tail_chunk_addr(base, r + 4, ch));
    const float2 x0 = unpack_bf16x2(q0.x), y0 = unpack_bf16x2(q0.y), z0 = unpack_bf16x2(q0.z), u0 = unpack_bf16x2(q0.w);
    const float2 x1 = unpack_bf16x2(q1.x), y1 = unpack_bf16x2(q1.y), z1 = unpack_bf16x2(q1.z), u1 = unpack_bf16x2(q1.w);
    acc[0] = fmaf(w0, x0.x, acc[0]), acc[1] = fmaf(w0, x0.y, acc[1]), acc[2] = fmaf(w0, y0.x, acc[2]), acc[3] = fmaf(w0, y0.y, acc[3]);
    acc[4] = fmaf(w0, z0.x, acc[4]), acc[5] = fmaf(w0, z0.y, acc[5]), acc[6] = fmaf(w0, u0.x, acc[6]), acc[7] = fmaf(w0, u0.y, acc[7]);
    b[0] = fmaf(w1, x1.x, b[0]), b[1] = fmaf(w1, x1.y, b[1]), b[2] = fmaf(w1, y1.x, b[2]), b[3] = fmaf(w1, y1.y, b[3]);
    b[4] = fmaf(w1, z1.x, b[4]), b[5] = fmaf(w1, z1.y, b[5]), b[6] = fmaf(w1, u1.x, b[6]), b[7] = fmaf(w1, u1.y, b[7]);
  }
  if (r < r1) {
    const float w0 = lds_f32(w_saddr + r * 4);
    const uint4 q0 = lds128(tail_chunk_addr(base, r, ch));
    const float2 x0 = unpack_bf16x2(q0.x), y0 = unpack_bf16x2(q0.y), z0 = unpack_bf16x2(q0.z), u0 = unpack_bf16x2(q0.w);
    acc[0] = fmaf(w0, x0.x, acc[0]), acc[1] = fmaf(w0, x0.y, acc[1]), acc[2] = fmaf(w0, y0.x, acc[2]), acc[3] = fmaf(w0, y0.y, acc[3]);
    acc[4] = fmaf(w0, z0.x, acc[4]), acc[5] = fmaf(w0, z0.y, acc[5]), acc[6] = fmaf(w0, u0.x, acc[6]), acc[7] = fmaf(w0, u0.y, acc[7]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] += b[j];
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
  }
}

// lanes 0..7 write their 8 values (scaled) as one 128-byte bf16 row
__device__ __forceinline__ void tail_store_row64(__nv_bfloat16* g, int lane, const float (&acc)[8], float scale) {
  if (lane < 8) {
    uint4 q;
    q.x = pack_bf16x2(acc[0] * scale, acc[1] * scale), q.y = pack_bf16x2(acc[2] * scale, acc[3] * scale);
    q.z = pack_bf16x2(acc[4] * scale, acc[5] * scale), q.w = pack_bf16x2(acc[6] * scale, acc[7] * scale);
    reinterpret_cast<uint4*>(g)[lane] = q;
  }
}

// The tail team: warps 2 and 3 of the CTA (64 threads; warp 2 is the TMEM allocator, idle in the main loop).
// Thread t = 0..63 of the team owns the rows t + 64 j, j < kTailSlots.
constexpr int kTailSlots = 5;                   // 64 * 5 = 320 >= 272 resident rows
constexpr int kTailW = 64 * kTailSlots;         // floats per weight array in shared memory
__device__ __forceinline__ void tail_team_sync() { asm volatile("bar.sync 2, 64;" ::: "memory"); }

}  // namespace missm
