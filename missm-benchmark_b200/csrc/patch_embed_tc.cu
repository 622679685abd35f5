// Patch embedding as an IMPLICIT GEMM on tcgen05 (forward): the stride = kernel Conv2d of
// CLIPVisionEmbeddings (languagebind/video/modeling_video.py:29-35, 45; image/modeling_image.py:184-196)
// followed by "+ position embedding" and the scatter into token order (row 0 of every image is the CLS row),
// without ever materialising the im2col matrix.
//
//   tok[img * (P + 1) + 1 + patch, :] = sum_k pixels(img, patch, k) * W[:, k]  +  pos[1 + patch, :]
//   k = (c * ps + i) * ps + j  (Conv2d weight order),  pixels(img, patch, k) = px[b][c][t][py*ps + i][pxi*ps + j]
//
// TMA cannot gather this operand: the image is fp32 NCHW and a patch row is 14 floats = 56 bytes, not a
// multiple of the 16-byte stride granularity of a tensor map (and im2col-mode maps want channels-last).  So
// the A operand is produced by the CTA itself: all 8 warps read the pixel rows of a 128-patch tile with
// coalesced 8-byte loads (a warp covers 2 x 896 contiguous bytes of one image row), convert to bf16 and store
// straight into the canonical 128B-swizzled K-major layout `tcgen05.mma` expects -- the whole [128 x Kpad] A
// tile (Kpad = 640: 160 KB) stays resident in shared memory and is reused by all D / 128 column tiles, whose
// weight k-blocks stream in by TMA through a 3-stage ring.  Accumulators are double-buffered in TMEM, the
// epilogue (4 warps) is the EPI_PATCH one of the GEMM kernels (gemm_common.cuh).
//
// Work per launch: 2 * (n_img * P) * D * K flop; bytes: pixels once (fp32), tokens once (fp32), weights + position
// table from L2.  The explicit path it replaces wrote and re-read a bf16 [n_img * P, Kpad] im2col matrix.
#include <cstdlib>
#include <cstring>

#define MISSM_KERNEL_TAG "patch_embed_tc"
#include "../../include/missm_b200.h"
#include "gemm_common.cuh"
#include "missm_common.cuh"

namespace missm {

constexpr int PE_THREADS = 256;      // warp 0 TMA (weights), 1 MMA, 2 TMEM alloc, 4-7 epilogue; all 8 fill A
constexpr int PE_BN = 128;
constexpr int PE_MAX_KB = 10;        // Kpad <= 640
constexpr int PE_STAGES = 3;
constexpr int PE_B_STAGE = PE_BN * BK * 2;
constexpr int PE_SMEM = PE_MAX_KB * A_STAGE_BYTES + PE_STAGES * PE_B_STAGE + 256 + 4 * kEpiStageBytes + 1024;

struct PatchEmbedParams {
  const float* px;
  const int32_t* sample_index;
  int C, T, H, W, gw, P, K, nkb;     // P = patches per image, K = C * ps * ps, nkb = Kpad / 64
  int num_mt;                        // 128-row tiles of the [n_img * P] patch list
  GemmParams g;                      // M = n_img * P, N = D, C = tok, aux_in = pos, patch_P = P
};

__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

template <int PS>
__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_tc_kernel(const __grid_constant__ CUtensorMap tmB, const PatchEmbedParams pp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                                        // [k-block][128 rows x 128 B], SW128 K-major
  uint8_t* sB = sA + PE_MAX_KB * A_STAGE_BYTES;              // [stage][128 rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + PE_STAGES * PE_B_STAGE);
  uint64_t* empty_bar = full_bar + PE_STAGES;
  uint64_t* tfull_bar = empty_bar + PE_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* sStage = reinterpret_cast<uint8_t*>(full_bar) + 256;

  const GemmParams& p = pp.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n_blk = p.N / PE_BN;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmB);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < PE_STAGES; ++i) mbar_init(&full_bar[i], 1), mbar_init(&empty_bar[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&tfull_bar[i], 1), mbar_init(&tempty_bar[i], 4);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * PE_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sA_u = smem_u32(sA);

  // ring / accumulator state lives across the M tiles of this CTA (every role advances only the counters it owns)
  int stage = 0, acc = 0;
  uint32_t phase = 0, acc_phase = 0;

  for (int mt = blockIdx.x; mt < pp.num_mt; mt += gridDim.x) {
    const int m0 = mt * BM;
    // ---- A tile: zero the k-blocks that contain padding columns, then gather the pixels ----------------
    const int kb_pad0 = pp.K / BK;                           // first k-block with columns >= K
    for (int i = threadIdx.x; i < (pp.nkb - kb_pad0) * (A_STAGE_BYTES / 16); i += PE_THREADS)
      sts128(sA_u + kb_pad0 * A_STAGE_BYTES + i * 16, make_uint4(0u, 0u, 0u, 0u));
    __syncthreads();
    const int items = pp.C * PS * BM;                        // (channel, kernel row) x tile row
    // three work items per thread and iteration: 21 independent 8-byte loads in flight (the gather is latency-bound;
    // with one item per iteration a 128-patch tile took 21 DRAM round trips)
    constexpr int UN = 3;
    for (int w0 = threadIdx.x; w0 < items; w0 += UN * PE_THREADS) {
      float2 v[UN][PS / 2];
      int m_loc[UN], k0s[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int w = w0 + u * PE_THREADS;
        const int m_local = w & (BM - 1), ci = w >> 7;
        const int row = m0 + m_local;
        m_loc[u] = (w < items && row < p.M) ? m_local : -1;
        k0s[u] = ci * PS;
        if (m_loc[u] < 0) continue;
        const int img = row / pp.P, patch = row - img * pp.P;
        const int py = patch / pp.gw, pxi = patch - py * pp.gw;
        const int c = ci / PS, i = ci - c * PS;
        const int b = img / pp.T, t = img - b * pp.T;
        const long src_b = pp.sample_index ? pp.sample_index[b] : b;
        const float2* src = reinterpret_cast<const float2*>(
            pp.px + (((src_b * pp.C + c) * pp.T + t) * pp.H + (py * PS + i)) * static_cast<long>(pp.W) + pxi * PS);
#pragma unroll
        for (int j = 0; j < PS / 2; ++j) v[u][j] = __ldg(src + j);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (m_loc[u] < 0) continue;
        const uint32_t row_off = sA_u + m_loc[u] * 128;
#pragma unroll
        for (int j = 0; j < PS / 2; ++j) {
          const int k = k0s[u] + 2 * j, kb = k >> 6, kk = k & 63;
          sts32(row_off + kb * A_STAGE_BYTES + (((kk >> 3) ^ (m_loc[u] & 7)) << 4) + (kk & 7) * 2,
                pack_bf16x2(v[u][j].x, v[u][j].y));
        }
      }
    }
    fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();

    if (warp == 0) {
      // ================================ TMA producer: weight k-blocks ====================
      for (int n_blk = 0; n_blk < num_n_blk; ++n_blk) {
        for (int kb = 0; kb < pp.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[stage], PE_B_STAGE);
            tma_load_2d(sB + stage * PE_B_STAGE, &tmB, &full_bar[stage], kb * BK, n_blk * PE_BN);
          }
          __syncwarp();
          if (++stage == PE_STAGES) stage = 0, phase ^= 1;
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ======================================
      const uint32_t idesc = umma_idesc_bf16_f32(BM, PE_BN, 0, 0);
      const uint64_t a_desc0 = umma_smem_desc_sw128(sA_u, 16, 1024);
      const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(sB), 16, 1024);
      for (int n_blk = 0; n_blk < num_n_blk; ++n_blk) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PE_BN;
        for (int kb = 0; kb < pp.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t a_desc = a_desc0 + kb * (A_STAGE_BYTES >> 4);
            const uint64_t b_desc = b_desc0 + stage * (PE_B_STAGE >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss(d_tmem, a_desc + k * 2, b_desc + k * 2, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == pp.nkb - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == PE_STAGES) stage = 0, phase ^= 1;
        }
        if (++acc == 2) acc = 0, acc_phase ^= 1;
      }
    } else if (warp >= 4) {
      // ================================ epilogue: + position row, scatter to token order =
      const int q = warp & 3;
      const uint32_t stg = smem_u32(sStage) + (warp - 4) * kEpiStageBytes;
      const int row0 = m0 + q * 32;
      for (int n_blk = 0; n_blk < num_n_blk; ++n_blk) {
        uint4 aux[8], aux_next[8];
        epilogue_aux_load<MISSM_EPI_PATCH>(p, row0, n_blk * PE_BN, lane, aux);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * PE_BN;
#pragma unroll 1
        for (int c = 0; c < PE_BN / 32; ++c) {
          const int col0 = n_blk * PE_BN + c * 32;
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          if (c + 1 < PE_BN / 32) epilogue_aux_load<MISSM_EPI_PATCH>(p, row0, col0 + 32, lane, aux_next);
          tmem_ld_wait();
          if (row0 < p.M) epilogue_chunk<MISSM_EPI_PATCH, true>(p, row0, col0, r, aux, stg, lane);   // warp-uniform
#pragma unroll
          for (int i = 0; i < 8; ++i) aux[i] = aux_next[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) acc = 0, acc_phase ^= 1;
      }
    }
    // the epilogue warps get here only after the last accumulator of this tile is complete, i.e. after every
    // MMA that reads sA has retired: the next tile may overwrite sA
    tc_fence_before();
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * PE_BN);
  }
}

}  // namespace missm

using namespace missm;

// returns 0 if launched, -1 if the shape is not handled (the caller then takes the explicit im2col + GEMM path)
extern "C" int missm_patch_embed_implicit(const float* pixels, const int32_t* sample_index, const void* w_bf16,
                                          const float* pos, float* tok, int32_t Bn, int32_t C, int32_t T, int32_t H,
                                          int32_t W, int32_t ps, int32_t Kpad, int32_t D, void* stream) {
  if (Bn == 0) return 0;
  const int K = C * ps * ps;
  static const bool off = getenv("MISSM_PATCH_EXPLICIT") != nullptr;   // A/B switch
  if (off || ps != 14 || Kpad % BK != 0 || Kpad > PE_MAX_KB * BK || Kpad < K || D % PE_BN != 0 || W % 2 != 0 || H % ps != 0 ||
      W % ps != 0 || T < 1)
    return -1;
  MISSM_REQUIRE(pixels && w_bf16 && pos && tok, "patch_embed_implicit: null pointer");
  MISSM_REQUIRE(reinterpret_cast<uintptr_t>(pixels) % 8 == 0, "patch_embed_implicit: pixels must be 8-byte aligned");
  PatchEmbedParams pp;
  memset(&pp, 0, sizeof(pp));
  const int gh = H / ps, gw = W / ps;
  pp.px = pixels, pp.sample_index = sample_index;
  pp.C = C, pp.T = T, pp.H = H, pp.W = W, pp.gw = gw, pp.P = gh * gw, pp.K = K, pp.nkb = Kpad / BK;
  const long M = static_cast<long>(Bn) * T * pp.P;
  MISSM_REQUIRE(M < (1l << 31), "patch_embed_implicit: too many patches");
  pp.num_mt = static_cast<int>((M + BM - 1) / BM);
  GemmParams& g = pp.g;
  g.M = static_cast<int>(M), g.N = D, g.K = Kpad;
  g.C = tok, g.ldc = D;
  g.aux_in = pos, g.ld_aux_in = D, g.patch_P = pp.P;
  g.col_scale = 1.f;
  CUtensorMap tmB;
  if (int rc = make_tmap_2d_bf16(&tmB, w_bf16, Kpad, D, Kpad, BK, PE_BN)) return rc;
  static bool configured = false;
  if (!configured) {
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_tc_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
    configured = true;
  }
  const int grid = pp.num_mt < persistent_sms() ? pp.num_mt : persistent_sms();
  patch_embed_tc_kernel<14><<<grid, PE_THREADS, PE_SMEM, static_cast<cudaStream_t>(stream)>>>(tmB, pp); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
