// Fused attention FORWARD on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), head_dim 64,
// non-causal, unmasked, whole key range resident (N <= 272): the ViT-L/14 spatial attention of the
// image / depth / thermal / video towers (N = 257), i.e. transformers 4.3x CLIPAttention's
// bmm -> softmax -> bmm chain called at languagebind/image/modeling_image.py:140.
//
// One persistent CTA per SM walks (sequence, head) items.  K and V of an item (all keys, bf16) are
// TMA-loaded once into 128B-swizzled shared memory (double-buffered across items), Q per 128-row
// tile (double-buffered).  Per query tile:
//   MMA warp     : S[128 x Nk] = Q K^T          tcgen05.mma, fp32 in TMEM columns [0, 272)
//   8 softmax warps (two per TMEM lane quarter, each owning half of the key columns of its 32
//                  rows): exact two-pass softmax straight from TMEM -- pass 1 row max (halves
//                  exchanged through shared memory), pass 2 exp2 + row sum, P written as packed
//                  bf16 to a SEPARATE TMEM region [272, 408) so that S is free again at once
//   MMA warp     : first S of the NEXT tile (overlaps the softmax tail / P V), then
//                  O[128 x 64] = P V           A operand from TMEM, V as MN-major smem operand
//   softmax warps: O / rowsum -> bf16 -> HBM and lse, interleaved between pass 1 and pass 2 of the
//                  next tile (that is when P V has finished)
// Scores never leave the SM; nothing is transposed in memory.
//
// Bound: MUFU (one ex2 per score) + TMEM read bandwidth.  Work per item = 4 * N^2 * 64 flop.
#include <cstdlib>

#define MISSM_KERNEL_TAG "attn_fwd_tc"
#include "../../include/missm_b200.h"
#include "attention_tail.cuh"
#include "missm_common.cuh"

namespace missm {

constexpr int FW_THREADS = 384;            // warp 0 TMA, 1 MMA, 2 TMEM alloc, 2-3 odd-row tail, 4-11 softmax
constexpr int FW_MAXN = 272;
constexpr int FW_KV_BYTES = FW_MAXN * 128; // one resident operand
constexpr int FW_TILE_BYTES = 128 * 128;
constexpr int FW_S_COL = 0, FW_P_COL = 272, FW_O_COL = 448;
constexpr float kLog2eFw = 1.4426950408889634f;

struct AttnFwdTcParams {
  int N, H, D, n_items;
  int sw;                 // N rounded up to 16
  int nt;                 // query tiles per item
  int tail;               // 1: row N-1 is computed by warps 2-3 on the CUDA cores (attention_tail.cuh), tiles cover [0, N-1)
  const __nv_bfloat16* qkv;
  long ld_qkv;
  __nv_bfloat16* out;
  long ld_o;
  float* lse;             // [n_seq, H, N] or null
  long long* trace;       // debugging only (MISSM_ATTN_TRACE_FWD): (tag, tile, clock) events of CTA 0
};

#ifdef MISSM_ATTN_TRACE_BUILD   // timeline tracing is compiled out of the production kernels (registers)
__device__ __forceinline__ void fw_trace(const AttnFwdTcParams& p, int slot, uint32_t& cnt, int tag, uint32_t n) {
  if (p.trace != nullptr && blockIdx.x == 0 && cnt < 330) {
    long long* t = p.trace + (slot * 330 + cnt) * 3;
    t[0] = tag, t[1] = n, t[2] = clock64();
    ++cnt;
  }
}
#else
__device__ __forceinline__ void fw_trace(const AttnFwdTcParams&, int, uint32_t&, int, uint32_t) {}
#endif

struct AttnFwdSmem {
  uint64_t kv_full[2], kv_empty[2];
  uint64_t q_full[2], q_empty[2];
  uint64_t s_full, p_full, o_full, o_empty;
  uint32_t tmem_base;
  float xmax[2][128];     // row max halves exchanged between the two warps of a row
  float xsum[2][2][128];  // row sum halves, [tile parity][column half][row]: read one tile later, in the epilogue
  float tail_w[kTailW];         // odd row: P
  float tail_x[4];              //          max / sum halves of the two tail warps
  float tail_part[64];          //          partial output of warp 3
};

__device__ __forceinline__ void fw_named_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool CO>
__global__ void MISSM_PERSISTENT_BOUNDS(CO)   // FW_THREADS threads, one CTA per SM
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16,
                   const AttnFwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sKV = smem;                              // [item buffer][K | V]: 2 x 2 x 34 KB
  uint8_t* sQ = sKV + 4 * FW_KV_BYTES;              // 2 x 16 KB
  uint8_t* sStage = sQ + 2 * FW_TILE_BYTES;           // 8 softmax warps x 2 KB output staging
  AttnFwdSmem* sh = reinterpret_cast<AttnFwdSmem*>(sStage + 8 * 2048);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm128), tma_prefetch_desc(&tm16);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh->kv_full[i], 1), mbar_init(&sh->kv_empty[i], p.tail ? 3 : 1);   // MMA commit (+ the two tail warps)
      mbar_init(&sh->q_full[i], 1), mbar_init(&sh->q_empty[i], 1);
    }
    mbar_init(&sh->s_full, 1), mbar_init(&sh->o_full, 1);
    mbar_init(&sh->p_full, 256), mbar_init(&sh->o_empty, 256);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&sh->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  // register pool of the CTA = 384 x kCoResidentRegs: control warpgroup (TMA, MMA, the two odd-row warps) 104, the two
  // softmax warpgroups 184 (128 x 104 + 256 x 184 <= 384 x 160)
  const int n_full = p.sw / 128, n_rem16 = (p.sw % 128) / 16;

  if (warp < 4) {
  if constexpr (CO) reg_dealloc<104>();
  if (warp == 0) {
    // ================================ TMA producer ====================================
    uint32_t it = 0, tcount = 0;
    const uint32_t kv_tx = 2u * static_cast<uint32_t>(n_full * FW_TILE_BYTES + n_rem16 * 16 * 128);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int s = item / p.H, h = item % p.H;
      const int kb = it & 1;
      mbar_wait(&sh->kv_empty[kb], ((it >> 1) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&sh->kv_full[kb], kv_tx);
        uint8_t* k_dst = sKV + (kb * 2 + 0) * FW_KV_BYTES;
        uint8_t* v_dst = sKV + (kb * 2 + 1) * FW_KV_BYTES;
        for (int j = 0; j < n_full; ++j) {
          tma_load_3d(k_dst + j * FW_TILE_BYTES, &tm128, &sh->kv_full[kb], p.D + h * 64, j * 128, s);
          tma_load_3d(v_dst + j * FW_TILE_BYTES, &tm128, &sh->kv_full[kb], 2 * p.D + h * 64, j * 128, s);
        }
        for (int j = 0; j < n_rem16; ++j) {
          const int row = n_full * 128 + j * 16;
          tma_load_3d(k_dst + row * 128, &tm16, &sh->kv_full[kb], p.D + h * 64, row, s);
          tma_load_3d(v_dst + row * 128, &tm16, &sh->kv_full[kb], 2 * p.D + h * 64, row, s);
        }
      }
      __syncwarp();
      for (int t = 0; t < p.nt; ++t, ++tcount) {
        const int qb = tcount & 1;
        mbar_wait(&sh->q_empty[qb], ((tcount >> 1) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&sh->q_full[qb], FW_TILE_BYTES);
          tma_load_3d(sQ + qb * FW_TILE_BYTES, &tm128, &sh->q_full[qb], h * 64, t * 128, s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ======================================
    const int n1 = p.sw > 256 ? 256 : p.sw;   // first S MMA width
    const int n2 = p.sw - n1;                 // second (0 or 16)
    const uint32_t idesc_s1 = umma_idesc_bf16_f32(128, n1, 0, 0);
    const uint32_t idesc_s2 = umma_idesc_bf16_f32(128, n2 > 0 ? n2 : 16, 0, 0);
    const uint32_t idesc_pv = umma_idesc_bf16_f32(128, 64, 0, 1);   // A from TMEM, B = V MN-major
    const uint64_t desc_kv = umma_smem_desc_sw128(smem_u32(sKV), 16, 1024);
    const uint64_t desc_q = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const int ksteps = p.sw / 16;
    const int my_items = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const uint32_t total_tiles = static_cast<uint32_t>(my_items * p.nt);
    uint32_t tr = 0;

    // S of local tile tc (local item index it = tc / nt)
    auto issue_s = [&](uint32_t tc) {
      const uint32_t it = tc / p.nt;
      if (tc % p.nt == 0) mbar_wait(&sh->kv_full[it & 1], (it >> 1) & 1);
      mbar_wait(&sh->q_full[tc & 1], (tc >> 1) & 1);
      tc_fence_after();
      if (lane == 0) fw_trace(p, 0, tr, 1, tc);
      if (elect_one_sync()) {
        const uint64_t qd = desc_q + (tc & 1) * (FW_TILE_BYTES >> 4);
        const uint64_t kd = desc_kv + ((it & 1) * 2 + 0) * (FW_KV_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_f16_ss(tmem + FW_S_COL, qd + k * 2, kd + k * 2, idesc_s1, k > 0);
          if (n2 > 0) umma_f16_ss(tmem + FW_S_COL + 256, qd + k * 2, kd + ((256 * 128) >> 4) + k * 2, idesc_s2, k > 0);
        }
        umma_commit(&sh->s_full);
        umma_commit(&sh->q_empty[tc & 1]);
      }
      __syncwarp();
      if (lane == 0) fw_trace(p, 0, tr, 2, tc);
    };

    if (total_tiles > 0) issue_s(0);
    for (uint32_t tc = 0; tc < total_tiles; ++tc) {
      const uint32_t it = tc / p.nt;
      mbar_wait(&sh->p_full, tc & 1);   // P(tc) is in TMEM and every read of S(tc) has finished
      tc_fence_after();
      if (lane == 0) fw_trace(p, 0, tr, 3, tc);
      if (tc + 1 < total_tiles) issue_s(tc + 1);
      mbar_wait(&sh->o_empty, (tc & 1) ^ 1);   // O(tc-1) has been read out (long done by now)
      tc_fence_after();
      if (lane == 0) fw_trace(p, 0, tr, 4, tc);
      if (elect_one_sync()) {
        const uint64_t vd = desc_kv + ((it & 1) * 2 + 1) * (FW_KV_BYTES >> 4);
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k)
          umma_f16_ts(tmem + FW_O_COL, tmem + FW_P_COL + k * 8, vd + k * 128, idesc_pv, k > 0);
        umma_commit(&sh->o_full);
        if (tc % p.nt == static_cast<uint32_t>(p.nt) - 1) umma_commit(&sh->kv_empty[it & 1]);
      }
      __syncwarp();
      if (lane == 0) fw_trace(p, 0, tr, 5, tc);
    }
  } else {
    // ========================= the odd row N-1 on the CUDA cores (warps 2 + 3) =========
    if (p.tail) {
      const int tw = warp - 2, t = tw * 32 + lane;
      const uint32_t w_s = smem_u32(sh->tail_w);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int s = item / p.H, h = item % p.H, kb = it & 1;
        const long row = static_cast<long>(s) * p.N + (p.N - 1);
        float a[64];
        tail_load_row64(p.qkv + row * p.ld_qkv + h * 64, a);       // q (pre-scaled)
        mbar_wait(&sh->kv_full[kb], (it >> 1) & 1);
        const uint32_t k_s = smem_u32(sKV + (kb * 2 + 0) * FW_KV_BYTES);
        const uint32_t v_s = smem_u32(sKV + (kb * 2 + 1) * FW_KV_BYTES);
        float sc[kTailSlots];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kTailSlots; ++j) {
          const int r = t + 64 * j;
          sc[j] = r < p.N ? tail_dot64(k_s, r, a) : -INFINITY;
          mx = fmaxf(mx, sc[j]);
        }
        mx = warp_max(mx);
        if (lane == 0) sh->tail_x[tw] = mx;
        tail_team_sync();
        mx = fmaxf(sh->tail_x[0], sh->tail_x[1]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kTailSlots; ++j) {
          const int r = t + 64 * j;
          const float e = r < p.N ? fast_ex2((sc[j] - mx) * kLog2eFw) : 0.f;
          sum += e;
          sts_f32(w_s + r * 4, e);
        }
        sum = warp_sum(sum);
        if (lane == 0) sh->tail_x[2 + tw] = sum;
        tail_team_sync();                                // weights and partial sums are visible
        sum = sh->tail_x[2] + sh->tail_x[3];
        float acc[8];
        tail_weighted_rowsum(v_s, w_s, tw == 0 ? 0 : 128, tw == 0 ? 128 : p.N, lane, acc);
        __syncwarp();                                    // every lane is done with K, V
        if (lane == 0) mbar_arrive(&sh->kv_empty[kb]);
        if (tw == 1 && lane < 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) sh->tail_part[lane * 8 + j] = acc[j];
        }
        tail_team_sync();
        if (tw == 0) {
          if (lane < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += sh->tail_part[lane * 8 + j];
          }
          tail_store_row64(p.out + row * p.ld_o + h * 64, lane, acc, 1.0f / sum);
          if (lane == 0 && p.lse != nullptr) p.lse[static_cast<long>(item) * p.N + p.N - 1] = mx + __logf(sum);
        }
      }
    }
  }
  } else {
    if constexpr (CO) reg_alloc<184>();
    // ========================= softmax + output ========================================
    const int g = (warp - 4) >> 2;       // column half
    const int q = warp & 3;              // TMEM lane quarter
    const int rloc = q * 32 + lane;      // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int ng = p.sw / 16;                              // 16-column groups of a row
    const int g_lo = g == 0 ? 0 : (ng + 1) / 2;            // this warp's groups [g_lo, g_hi)
    const int g_hi = g == 0 ? (ng + 1) / 2 : ng;
    const uint32_t s_addr = tmem + lane_addr + FW_S_COL;
    const uint32_t p_addr = tmem + lane_addr + FW_P_COL;
    uint32_t tr = 0;
    const bool tracer = (q == 0 && lane == 0);
    const uint32_t stage = smem_u32(sStage) + (warp - 4) * 2048;

    // epilogue of a finished tile: O / l -> bf16 (this warp writes columns [g*32, g*32+32)), lse
    auto epilogue = [&](uint32_t tc, int item, int tile, float mx, float sum_own, bool has_rows) {
      const float sum = sum_own + sh->xsum[tc & 1][g ^ 1][rloc];
      mbar_wait(&sh->o_full, tc & 1);
      tc_fence_after();
      if (tracer) fw_trace(p, 1 + g, tr, 13, tc);
      uint32_t o[32];
      if (has_rows) {
        tmem_ld_32x32b_x32(tmem + lane_addr + FW_O_COL + g * 32, o);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&sh->o_empty);
      const int row = tile * 128 + rloc;
      if (has_rows) {     // warp-uniform
        const int s = item / p.H, h = item % p.H;
        const float inv = 1.0f / sum;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(o[j]) * inv;
        const int wrow0 = tile * 128 + q * 32;
        warp_store_tile32_bf16(stage, lane, v, p.out + (static_cast<long>(s) * p.N + wrow0) * p.ld_o + h * 64 + g * 32,
                               p.ld_o, p.N - wrow0);
        if (row < p.N && g == 0 && p.lse != nullptr) p.lse[static_cast<long>(item) * p.N + row] = mx + __logf(sum);
      }
      if (tracer) fw_trace(p, 1 + g, tr, 14, tc);
    };

    uint32_t tc = 0;
    int prev_item = -1, prev_tile = 0;
    float prev_mx = 0.f, prev_sum = 0.f;
    bool prev_rows = false;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int t = 0; t < p.nt; ++t, ++tc) {
        const bool has_rows = t * 128 + q * 32 < p.N;      // warp-uniform
        if (tracer) fw_trace(p, 1 + g, tr, 10, tc);
        mbar_wait(&sh->s_full, tc & 1);
        tc_fence_after();
        if (tracer) fw_trace(p, 1 + g, tr, 11, tc);
        // ---- pass 1: row maximum over this warp's valid key columns
        float mx = -INFINITY;
        if (has_rows) {
          for (int gi = g_lo; gi < g_hi; gi += 2) {
            if (gi + 1 < g_hi) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(s_addr + gi * 16, r);
              tmem_ld_wait();
              if (gi * 16 + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (gi * 16 + j < p.N) mx = fmaxf(mx, __uint_as_float(r[j]));
              }
            } else {
              uint32_t r[16];
              tmem_ld_32x32b_x16(s_addr + gi * 16, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (gi * 16 + j < p.N) mx = fmaxf(mx, __uint_as_float(r[j]));
            }
          }
        }
        sh->xmax[g][rloc] = mx;
        if (tracer) fw_trace(p, 1 + g, tr, 12, tc);
        fw_named_sync();                        // (also orders the xsum of tile tc-1, written before its p_full arrive)
        mx = fmaxf(mx, sh->xmax[g ^ 1][rloc]);
        if (tracer) fw_trace(p, 1 + g, tr, 15, tc);
        // ---- pass 2: P = exp2(S*log2e - max*log2e) -> bf16 into the P region; partial row sum.
        //      The P region is free once P V of the previous tile has completed (o_full).  The epilogue below waits for
        //      the same phase again, which returns at once: phase tc of o_full cannot complete before every softmax
        //      thread has arrived on o_empty, i.e. has passed that second wait.
        if (prev_item >= 0) {
          mbar_wait(&sh->o_full, (tc - 1) & 1);
          tc_fence_after();
        }
        const float mx2 = mx * kLog2eFw;
        float sum = 0.f;
        if (has_rows) {
          for (int gi = g_lo; gi < g_hi; gi += 2) {
            if (gi + 1 < g_hi) {
              uint32_t r[32], pk[16];
              tmem_ld_32x32b_x32(s_addr + gi * 16, r);
              tmem_ld_wait();
              const bool full = gi * 16 + 32 <= p.N;
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float a = fast_ex2(fmaf(__uint_as_float(r[j]), kLog2eFw, -mx2));
                float b = fast_ex2(fmaf(__uint_as_float(r[j + 1]), kLog2eFw, -mx2));
                if (!full) {
                  if (gi * 16 + j >= p.N) a = 0.f;
                  if (gi * 16 + j + 1 >= p.N) b = 0.f;
                }
                sum += a + b;
                pk[j >> 1] = pack_bf16x2(a, b);
              }
              tmem_st_32x32b_x16(p_addr + gi * 8, pk);
            } else {
              uint32_t r[16], pk[8];
              tmem_ld_32x32b_x16(s_addr + gi * 16, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float a = (gi * 16 + j < p.N) ? fast_ex2(fmaf(__uint_as_float(r[j]), kLog2eFw, -mx2)) : 0.f;
                float b = (gi * 16 + j + 1 < p.N) ? fast_ex2(fmaf(__uint_as_float(r[j + 1]), kLog2eFw, -mx2)) : 0.f;
                sum += a + b;
                pk[j >> 1] = pack_bf16x2(a, b);
              }
              tmem_st_32x32b_x8(p_addr + gi * 8, pk);
            }
          }
          tmem_st_wait();
        }
        sh->xsum[tc & 1][g][rloc] = sum;       // read by the other half's warp in the epilogue of this tile, one bar.sync later
        tc_fence_before();
        mbar_arrive(&sh->p_full);
        if (tracer) fw_trace(p, 1 + g, tr, 16, tc);
        // ---- previous tile's output while the MMA warp computes S of the next tile (S is single-buffered: the
        //      softmax warps used to sit idle here for ~1.6 k of every ~8 k cycles).  Its P V finished long ago; the
        //      other half's row sum was published before that warp's p_full arrive of tile tc-1 and ordered by the
        //      bar.sync of this iteration.
        if (prev_item >= 0) epilogue(tc - 1, prev_item, prev_tile, prev_mx, prev_sum, prev_rows);
        prev_item = item, prev_tile = t, prev_mx = mx, prev_sum = sum, prev_rows = has_rows;
      }
    }
    // ---- the last tile of this CTA
    if (prev_item >= 0) {
      fw_named_sync();
      epilogue(tc - 1, prev_item, prev_tile, prev_mx, prev_sum, prev_rows);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// returns 0 if launched, -1 if the shape is not handled by this kernel (caller falls back to the
// general mma.sync path), > 0 on error
int attention_fwd_tc(const missm_attn_args* a, cudaStream_t stream) {
  const bool ok = !a->causal && a->key_mask == nullptr && a->s_in == 1 && a->tok_stride == 1 &&
                  a->seq_outer == a->N && a->N <= FW_MAXN && a->N >= 16 && a->head_dim == 64;
  if (!ok) return -1;
  CUtensorMap tm128, tm16;
  const uint64_t seq = static_cast<uint64_t>(a->N) * a->ld_qkv;
  if (int rc = make_tmap_3d_bf16(&tm128, a->qkv, 3 * static_cast<uint64_t>(a->D), a->N, a->n_seq, a->ld_qkv, seq, 64, 128)) return rc;
  if (int rc = make_tmap_3d_bf16(&tm16, a->qkv, 3 * static_cast<uint64_t>(a->D), a->N, a->n_seq, a->ld_qkv, seq, 64, 16)) return rc;
  AttnFwdTcParams p;
  p.N = a->N, p.H = a->H, p.D = a->D, p.n_items = a->n_seq * a->H;
  p.sw = (a->N + 15) / 16 * 16;
  p.tail = attention_tail_enabled(a->N) ? 1 : 0;
  p.nt = p.tail ? a->N / 128 : (a->N + 127) / 128;
  if (p.tail && getenv("MISSM_ATTN_TAIL_SKIP") != nullptr) p.tail = 0;   // TIMING EXPERIMENT ONLY: row N-1 is left uncomputed
  p.qkv = static_cast<const __nv_bfloat16*>(a->qkv), p.ld_qkv = a->ld_qkv;
  p.out = static_cast<__nv_bfloat16*>(a->out), p.ld_o = a->ld_o, p.lse = a->lse;
  p.trace = nullptr;
  const int smem = 4 * FW_KV_BYTES + 2 * FW_TILE_BYTES + 8 * 2048 + static_cast<int>(sizeof(AttnFwdSmem)) + 1024;
  static bool configured = false;
  if (!configured) {
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int grid = p.n_items < persistent_sms() ? p.n_items : persistent_sms();
  const char* trace_path = getenv("MISSM_ATTN_TRACE_FWD");   // debugging aid (synchronises!)
  if (trace_path != nullptr) {
    const size_t nb = 3 * 330 * 3 * sizeof(long long);
    long long* d = nullptr;
    MISSM_CHECK_CUDA(cudaMalloc(&d, nb));
    MISSM_CHECK_CUDA(cudaMemsetAsync(d, 0, nb, stream));
    p.trace = d;
    attn_fwd_tc_kernel<false><<<grid, FW_THREADS, smem, stream>>>(tm128, tm16, p); note_launch();
    MISSM_CHECK_CUDA(cudaStreamSynchronize(stream));
    long long* h = static_cast<long long*>(malloc(nb));
    MISSM_CHECK_CUDA(cudaMemcpy(h, d, nb, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "w")) {
      for (size_t i = 0; i < nb / 24; ++i)
        if (h[3 * i] != 0) fprintf(f, "%zu %lld %lld %lld\n", i / 330, h[3 * i], h[3 * i + 1], h[3 * i + 2]);
      fclose(f);
    }
    free(h);
    cudaFree(d);
    return 0;
  }
  if (coresident())
    attn_fwd_tc_kernel<true><<<grid, FW_THREADS, smem, stream>>>(tm128, tm16, p);
  else
    attn_fwd_tc_kernel<false><<<grid, FW_THREADS, smem, stream>>>(tm128, tm16, p);
  note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace missm
