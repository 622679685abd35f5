// Missing-modality mask compaction.
//
// The reference runs every tower on the full batch and only overwrites the rows of missing
// samples in the fusion head (mask = `missing_index == missing_type_index[modal]`,
// src/model/baseline.py:57).  Here each tower receives only its PRESENT samples:
//     present_t = ascending { b : missing_index[b] != code_t }       (integer, bit-exact)
// which is what torch.nonzero(missing_index != code_t) returns.  One warp per tower: each lane
// tests a strip of 32 samples per iteration, __ballot_sync builds the 32-bit presence word, the
// popcount of the lower lanes is the exclusive prefix inside the strip, and a running base
// carries across strips -- no atomics, so the order is the ascending sample order.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

struct CompactCodes {
  int32_t code[MISSM_MAX_TOWERS];
};

__global__ void compact_mask_kernel(const int64_t* __restrict__ missing_index, int Bn,
                                    CompactCodes codes, int n_towers,
                                    int32_t* __restrict__ present_idx,   // [n_towers, Bn]
                                    int32_t* __restrict__ slot_of,      // [n_towers, Bn] or null
                                    int32_t* __restrict__ counts) {     // [n_towers]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= n_towers) return;
  const int code = codes.code[warp];
  int32_t* idx = present_idx + static_cast<long>(warp) * Bn;
  int32_t* slot = slot_of ? slot_of + static_cast<long>(warp) * Bn : nullptr;
  int base = 0;
  for (int b0 = 0; b0 < Bn; b0 += 32) {
    const int b = b0 + lane;
    const bool present = (b < Bn) && (missing_index[b] != static_cast<int64_t>(code));
    const unsigned word = __ballot_sync(0xffffffffu, present);
    const int pos = base + __popc(word & ((1u << lane) - 1u));
    if (present) idx[pos] = b;
    if (slot && b < Bn) slot[b] = present ? pos : -1;
    base += __popc(word);
  }
  if (lane == 0) counts[warp] = base;
}

// dst[b, :] = slot_of[b] >= 0 ? src[slot_of[b], :] : 0      (embeddings back to batch order,
// ZERO-filled for missing rows: regression / inter_attention multiply them by 0 and 0*NaN != 0)
__global__ void scatter_rows_zero_kernel(const float* __restrict__ src, const int32_t* __restrict__ slot_of,
                                         float* __restrict__ dst, int Bn, int P) {
  const int groups = P / 4;
  const long total = static_cast<long>(Bn) * groups;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = i / groups;
    const int c = static_cast<int>(i % groups) * 4;
    const int s = slot_of[b];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= 0) v = *reinterpret_cast<const float4*>(src + static_cast<long>(s) * P + c);
    *reinterpret_cast<float4*>(dst + b * P + c) = v;
  }
}

// dst[r, :] = src[idx[r], :]   rows of `row_bytes` bytes (multiple of 16)
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ idx,
                                   uint4* __restrict__ dst, int n_rows, long vec_per_row) {
  const long total = static_cast<long>(n_rows) * vec_per_row;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / vec_per_row, c = i % vec_per_row;
    dst[i] = src[static_cast<long>(idx[r]) * vec_per_row + c];
  }
}

}  // namespace missm

using namespace missm;

extern "C" int missm_compact_mask(const int64_t* missing_index, int32_t Bn, const int32_t* codes_host,
                                  int32_t n_towers, int32_t* present_idx, int32_t* slot_of,
                                  int32_t* counts, void* stream) {
  MISSM_REQUIRE(n_towers >= 1 && n_towers <= MISSM_MAX_TOWERS, "compact_mask: n_towers=%d", n_towers);
  CompactCodes cc;
  for (int i = 0; i < MISSM_MAX_TOWERS; ++i) cc.code[i] = i < n_towers ? codes_host[i] : -1;
  compact_mask_kernel<<<1, 32 * n_towers, 0, static_cast<cudaStream_t>(stream)>>>(
      missing_index, Bn, cc, n_towers, present_idx, slot_of, counts); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_scatter_rows_zero(const float* src, const int32_t* slot_of, float* dst,
                                       int32_t Bn, int32_t P, void* stream) {
  if (Bn == 0) return 0;
  MISSM_REQUIRE(P % 4 == 0, "scatter_rows_zero: P=%d", P);
  const long total = static_cast<long>(Bn) * (P / 4);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  scatter_rows_zero_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, slot_of, dst, Bn, P); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_gather_rows(const void* src, const int32_t* idx, void* dst, int32_t n_rows,
                                 int64_t row_bytes, void* stream) {
  if (n_rows == 0) return 0;
  MISSM_REQUIRE(row_bytes % 16 == 0, "gather_rows: row_bytes=%ld", (long)row_bytes);
  const long vec = row_bytes / 16, total = n_rows * vec;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
  gather_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), idx, static_cast<uint4*>(dst), n_rows, vec); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
