// Shared pieces of the tcgen05 GEMM kernels (1-CTA and CTA-pair): parameters and the fused
// epilogues applied to one accumulator row x 32 columns held in registers.
#pragma once
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kEpiWarps = 8;                         // two per TMEM lane quarter
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;  // 4 control warps + epilogue warps
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmParams {
  int M, N, K;
  int num_m_blk, num_n_blk, num_splits, kblk_per_split, num_kblk;
  int a_mn, b_mn;
  void* C;
  int ldc;
  const float* bias;
  float col_scale;
  int scale_cols;
  const void* aux_in;
  int ld_aux_in;
  void* aux_out;
  int ld_aux_out;
  int patch_P;
  int atomic_out;
  float* colsum_out;
  float* colsum_part;    // [ceil(M / 32), N]: per-warp partial column sums of the stored values
  int stream_k;          // 1: every CTA (pair) owns one contiguous slice of the linearised (tile, k-block) space
  long sk_per_cta;       //    k-blocks per slice
};

// The work list of one persistent CTA (pair).  Two schedules:
//   strided  : unit w = index + i * stride over tiles x uniform K splits (tile = w % tiles, split = w / tiles)
//   stream-K : one contiguous slice [index * sk_per_cta, ...) of the (tile-major, k-block-minor) sequence;
//              tiles cut by a slice boundary are accumulated with fp32 atomics into a zeroed C.  Used for
//              wgrad, whose output grid (48-64 pair tiles, K = 14 906) would leave 14-35 % of the SMs idle.
struct GemmWork {
  int tiles_mn, num_kblk, kblk_per_split, total_work, stride, w;
  long pos, end;
  bool sk;
  __device__ __forceinline__ GemmWork(const GemmParams& p, int index, int count) {
    tiles_mn = p.num_m_blk * p.num_n_blk, num_kblk = p.num_kblk, kblk_per_split = p.kblk_per_split;
    total_work = tiles_mn * p.num_splits, stride = count, w = index;
    sk = p.stream_k != 0;
    const long total = static_cast<long>(tiles_mn) * num_kblk;
    pos = index * p.sk_per_cta;
    end = pos + p.sk_per_cta < total ? pos + p.sk_per_cta : total;
  }
  // next unit: tile (n fastest: m_blk = tile / num_n_blk) and its k-block range [kb0, kb1)
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (sk) {
      if (pos >= end) return false;
      tile = static_cast<int>(pos / num_kblk);
      kb0 = static_cast<int>(pos % num_kblk);
      const long left = end - pos;
      kb1 = kb0 + left < num_kblk ? kb0 + static_cast<int>(left) : num_kblk;
      pos += kb1 - kb0;
      return true;
    }
    if (w >= total_work) return false;
    tile = w % tiles_mn;
    kb0 = (w / tiles_mn) * kblk_per_split;
    kb1 = kb0 + kblk_per_split < num_kblk ? kb0 + kblk_per_split : num_kblk;
    w += stride;
    return true;
  }
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ +
                                    256 /*barriers + tmem slot*/ + kEpiWarps * 32 * 128 /*epilogue staging*/;
};

// ---------------------------------------------------------------------------------------
// Epilogue of one warp's 32 accumulator rows x 32 consecutive columns.
//
// tcgen05.ld hands every thread ONE ROW of the accumulator, but a warp-wide global access in that
// layout touches 32 different 128-byte lines per instruction (measured: the row-per-thread stores
// made the GELU / dGELU GEMMs 1.45x slower than the plain one).  So every global tile goes through a
// private 4 KB shared-memory staging tile per warp (128-byte row pitch, 16-byte chunks XOR-swizzled
// by row & 7 -> conflict-free both ways):   global <-coalesced-> staging <-row per thread-> registers.
// A coalesced instruction covers 4 rows x 128 B (fp32) or 8 rows x 64 B (bf16).
// ---------------------------------------------------------------------------------------
constexpr int kEpiStageBytes = 32 * 128;   // per epilogue warp

__device__ __forceinline__ uint32_t stage_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

// global rows of a tile: local row r (0..31) -> global row index, or -1 (skip)
struct RowMapIdentity {
  int row0, M;
  __device__ __forceinline__ long operator()(int r) const { return row0 + r < M ? row0 + r : -1; }
};
struct RowMapPatchPos {   // position-embedding row of patch (row % P): 1 + row % P
  int row0, M, P;
  __device__ __forceinline__ long operator()(int r) const { return row0 + r < M ? 1 + (row0 + r) % P : -1; }
};
struct RowMapPatchOut {   // token row of (sample, patch): sample * (P + 1) + 1 + patch
  int row0, M, P;
  __device__ __forceinline__ long operator()(int r) const {
    const int row = row0 + r;
    return row < M ? static_cast<long>(row / P) * (P + 1) + 1 + row % P : -1;
  }
};

// coalesced global -> staging.  Two phases so the loads of a tile can be issued early.
template <bool F32, typename RowMap>
__device__ __forceinline__ void tile_ldg(const void* base, long ld, int col0, int ncols, const RowMap& rm, int lane,
                                         uint4 (&val)[8]) {
  constexpr int CH = F32 ? 8 : 4, RPI = 32 / CH, EPC = F32 ? 4 : 8, NI = 32 / RPI, ES = F32 ? 4 : 2;
  const int c = lane % CH;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int r = i * RPI + lane / CH;
    const long g = rm(r);
    val[i] = make_uint4(0u, 0u, 0u, 0u);
    if (g >= 0 && c * EPC < ncols)
      val[i] = *reinterpret_cast<const uint4*>(static_cast<const char*>(base) + (g * ld + col0 + c * EPC) * ES);
  }
}
template <bool F32>
__device__ __forceinline__ void tile_sts(uint32_t stage, int lane, const uint4 (&val)[8]) {
  constexpr int CH = F32 ? 8 : 4, RPI = 32 / CH, NI = 32 / RPI;
  const int c = lane % CH;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int r = i * RPI + lane / CH;
    sts128(stage + stage_off(r, c), val[i]);
  }
}
// staging -> this thread's row as 32 floats
template <bool F32>
__device__ __forceinline__ void tile_row_read(uint32_t stage, int lane, float (&x)[32]) {
  if constexpr (F32) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 v = lds128(stage + stage_off(lane, c));
      x[4 * c] = __uint_as_float(v.x), x[4 * c + 1] = __uint_as_float(v.y);
      x[4 * c + 2] = __uint_as_float(v.z), x[4 * c + 3] = __uint_as_float(v.w);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 q = lds128(stage + stage_off(lane, c));
      const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), d = unpack_bf16x2(q.z), e = unpack_bf16x2(q.w);
      x[8 * c] = a.x, x[8 * c + 1] = a.y, x[8 * c + 2] = b.x, x[8 * c + 3] = b.y;
      x[8 * c + 4] = d.x, x[8 * c + 5] = d.y, x[8 * c + 6] = e.x, x[8 * c + 7] = e.y;
    }
  }
}
// this thread's row (32 floats) -> staging as fp32 or bf16
template <bool F32>
__device__ __forceinline__ void tile_row_write(uint32_t stage, int lane, const float (&x)[32]) {
  if constexpr (F32) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128(stage + stage_off(lane, c), make_uint4(__float_as_uint(x[4 * c]), __float_as_uint(x[4 * c + 1]),
                                                    __float_as_uint(x[4 * c + 2]), __float_as_uint(x[4 * c + 3])));
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 q;
      q.x = pack_bf16x2(x[8 * c], x[8 * c + 1]), q.y = pack_bf16x2(x[8 * c + 2], x[8 * c + 3]);
      q.z = pack_bf16x2(x[8 * c + 4], x[8 * c + 5]), q.w = pack_bf16x2(x[8 * c + 6], x[8 * c + 7]);
      sts128(stage + stage_off(lane, c), q);
    }
  }
}
// staging -> global, coalesced; `atomic` (fp32 only): red.global.add instead of a store
template <bool F32, typename RowMap>
__device__ __forceinline__ void tile_stg(void* base, long ld, int col0, int ncols, const RowMap& rm, uint32_t stage,
                                         int lane, bool atomic) {
  constexpr int CH = F32 ? 8 : 4, RPI = 32 / CH, EPC = F32 ? 4 : 8, NI = 32 / RPI, ES = F32 ? 4 : 2;
  const int c = lane % CH;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int r = i * RPI + lane / CH;
    const long g = rm(r);
    if (g >= 0 && c * EPC < ncols) {
      const uint4 v = lds128(stage + stage_off(r, c));
      char* dst = static_cast<char*>(base) + (g * ld + col0 + c * EPC) * ES;
      if (F32 && atomic) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v.x)),
                     "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                     : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst) = v;
      }
    }
  }
}

// `row0`: global row of this warp's lane 0; `acc`: this thread's row (row0 + lane) of the fp32
// accumulator, columns [col0, col0 + 32).  All 32 lanes must call (warp-collective).
// The auxiliary input tile of a chunk (residual / pre-activation / position rows), as coalesced
// global loads into registers.  Issued one chunk AHEAD of its use (the epilogue warps have no other
// way to hide a global-memory round trip: two warps per scheduler, one chunk at a time).
template <int EPI>
__device__ __forceinline__ void epilogue_aux_load(const GemmParams& p, int row0, int col0, int lane, uint4 (&aux)[8]) {
  if (row0 >= p.M || col0 >= p.N) return;
  const int ncols = min(32, p.N - col0);
  if constexpr (EPI == MISSM_EPI_RESID) tile_ldg<true>(p.aux_in, p.ld_aux_in, col0, ncols, RowMapIdentity{row0, p.M}, lane, aux);
  if constexpr (EPI == MISSM_EPI_DGELU) tile_ldg<false>(p.aux_in, p.ld_aux_in, col0, ncols, RowMapIdentity{row0, p.M}, lane, aux);
  if constexpr (EPI == MISSM_EPI_PATCH)
    tile_ldg<true>(p.aux_in, p.ld_aux_in, col0, ncols, RowMapPatchPos{row0, p.M, p.patch_P}, lane, aux);
}
template <int EPI>
constexpr bool epilogue_has_aux() { return EPI == MISSM_EPI_RESID || EPI == MISSM_EPI_DGELU || EPI == MISSM_EPI_PATCH; }

template <int EPI, bool OUT_F32>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row0, int col0, const uint32_t (&acc)[32],
                                               const uint4 (&aux)[8], uint32_t stage, int lane) {
  const int ncols = min(32, p.N - col0);  // N % 8 == 0 is enforced by the host
  const RowMapIdentity rows{row0, p.M};
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);

  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols) {
        float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
      }
    }
  }

  if constexpr (EPI == MISSM_EPI_LINEAR) {
    if (p.scale_cols > 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.scale_cols) v[j] *= p.col_scale;
    }
  } else if constexpr (EPI == MISSM_EPI_GELU) {
    // pre-activation u (bf16) out first, then the activation
    tile_row_write<false>(stage, lane, v);
    __syncwarp();
    tile_stg<false>(p.aux_out, p.ld_aux_out, col0, ncols, rows, stage, lane, false);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
  } else if constexpr (EPI == MISSM_EPI_RESID || EPI == MISSM_EPI_PATCH) {
    float x[32];
    tile_sts<true>(stage, lane, aux);
    __syncwarp();
    tile_row_read<true>(stage, lane, x);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += x[j];
  } else if constexpr (EPI == MISSM_EPI_DGELU) {
    float x[32];
    tile_sts<false>(stage, lane, aux);
    __syncwarp();
    tile_row_read<false>(stage, lane, x);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= quick_gelu_grad(x[j]);
  }

  tile_row_write<OUT_F32>(stage, lane, v);
  if (p.colsum_out != nullptr) {   // warp-uniform
    if (row0 + lane >= p.M) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    const float cs = warp_colsum32(v, lane);
    if (lane < ncols) atomicAdd(p.colsum_out + col0 + lane, cs);
  }
  if (p.colsum_part != nullptr) {   // warp-uniform
    // column sums of this warp's 32 rows of the tile AS STORED (bf16-rounded when C is bf16), one coalesced 128-byte
    // store per chunk into the row of the partial buffer that this warp owns: bias gradients without a second pass
    // over C and without atomics
    float t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = v[j];
      if constexpr (!OUT_F32) x = __bfloat162float(__float2bfloat16(x));
      t[j] = (row0 + lane < p.M) ? x : 0.f;
    }
    const float cs = warp_colsum32(t, lane);
    if (lane < ncols) p.colsum_part[static_cast<long>(row0 >> 5) * p.N + col0 + lane] = cs;
  }
  __syncwarp();
  if constexpr (EPI == MISSM_EPI_PATCH)
    tile_stg<OUT_F32>(p.C, p.ldc, col0, ncols, RowMapPatchOut{row0, p.M, p.patch_P}, stage, lane, false);
  else
    tile_stg<OUT_F32>(p.C, p.ldc, col0, ncols, rows, stage, lane, p.atomic_out != 0);
  __syncwarp();
}

}  // namespace missm
