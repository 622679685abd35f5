// Fused attention BACKWARD on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), head_dim 64,
// non-causal, unmasked, N <= 272 tokens: the ViT-L/14 spatial attention of the image / depth /
// thermal / video towers (N = 257) -- the autograd twin of transformers 4.3x CLIPAttention's
// bmm -> softmax -> bmm chain called at languagebind/image/modeling_image.py:140.
//
// Two kernels from one template, neither of which ever puts a score in shared memory or HBM:
//   DKV (row index = key):    S^T = K Q^T, dP^T = V dO^T       (tcgen05.mma, smem x smem -> TMEM)
//                             P^T = exp2(S^T log2e - lse), dS^T = P^T o (dP^T - delta)
//                               -> bf16, written back IN PLACE in TMEM (tcgen05.ld / tcgen05.st)
//                             dV += P^T dO,  dK += dS^T Q      (tcgen05.mma, A operand from TMEM)
//   DQ  (row index = query):  S = Q K^T, dP = dO V^T;  dS likewise;  dQ += dS K
// One persistent CTA per SM walks (sequence, head) items.  The "column side" operands of an item
// (DKV: all of Q and dO; DQ: all of K and V; 272 rows x 64, 34 KB each) are TMA-loaded once into
// 128B-swizzled shared memory, double-buffered across items; the 128-row "row side" tiles are
// double-buffered across tiles.  Columns are processed in chunks of <= 96: the two fp32 score
// chunks (2 x 96 TMEM columns) are double-buffered, two softmax warpgroups ping-pong over the chunks
// (one thread per row, 32-column register blocks), and the MMA warp issues the score MMAs of chunk
// n+1 before the accumulation MMAs of chunk n, so tensor pipe, MUFU and TMEM traffic overlap.
// The same shared-memory tile serves as K-major operand of the score MMAs and as MN-major operand of
// the accumulation MMAs -- no transposed copy of anything exists.
//
// TMEM columns: [0,96) S_0  [96,192) dP_0  [192,288) S_1  [288,384) dP_1  [384,512) accumulators.
// Bound: MUFU (one ex2 per score and kernel) / tensor.  Algorithmic work per item: DKV 8, DQ 6
// (together the usual "2.5 x forward" counts 10) x N^2 x 64 flop.
#include <cstdlib>

#define MISSM_KERNEL_TAG "attn_bwd_tc"
#include "../../include/missm_b200.h"
#include "attention_tail.cuh"
#include "missm_common.cuh"

namespace missm {

constexpr int BW_THREADS = 384;      // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 column statistics, 2-3 odd-row tail, 4-7 / 8-11 softmax WGs
constexpr int BW_CW = 96;            // score chunk width (columns)
constexpr int BW_MAXN = 272;         // resident rows (multiple of 16)
constexpr int BW_RES_BYTES = BW_MAXN * 128;   // one resident operand (rows of 64 bf16)
constexpr int BW_TILE_BYTES = 128 * 128;      // one row-side tile
constexpr int BW_STAT = 288;         // floats per column-statistics array
constexpr int BW_ACC_COL = 384;
constexpr float kLog2eBw = 1.4426950408889634f;

struct AttnBwdTcParams {
  int N, H, D, n_items;
  int sw;      // N rounded up to 16
  int nt;      // row tiles (128 rows)
  int nc;      // column chunks
  int tail;    // DQ only, 1: query row N-1 is computed by warps 2-3 on the CUDA cores (tiles cover [0, N-1));
               // DKV gets nt = (N-1)/128 with tail = 0: its key row N-1 comes from attn_delta_tail_kernel
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* d_out;
  long ld_o;
  const float* lse;    // [n_seq, H, N]
  const float* delta;  // [n_seq, H, N]
  __nv_bfloat16* dqkv;
  long ld_qkv;
  float q_scale;
  long long* trace;    // debugging only (MISSM_ATTN_TRACE): [4096] x (tag, step, clock) of CTA 0
};

#ifdef MISSM_ATTN_TRACE_BUILD   // timeline tracing is compiled out of the production kernels (registers)
__device__ __forceinline__ void bw_trace(const AttnBwdTcParams& p, int slot_base, uint32_t& cnt, int tag, uint32_t n) {
  if (p.trace != nullptr && blockIdx.x == 0 && cnt < 330) {
    long long* t = p.trace + (slot_base * 330 + cnt) * 3;
    t[0] = tag, t[1] = n, t[2] = clock64();
    ++cnt;
  }
}
#else
__device__ __forceinline__ void bw_trace(const AttnBwdTcParams&, int, uint32_t&, int, uint32_t) {}
#endif

struct AttnBwdSmem {
  uint64_t res_full[2], res_empty[2];
  uint64_t a_full[2], a_empty[2];
  uint64_t s_full[2], p_full[2], sbuf_empty[2];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t stat_full[2], stat_empty[2];
  uint32_t tmem_base;
  float tail_w[2][kTailW];   // odd row: P | dS (DQ: dS | partial dQ of warp 3)
};

// position in the flattened (item, row tile, column chunk) sequence of this CTA
struct BwCursor {
  int item, tile, chunk;
  uint32_t n, tcount, it;
  __device__ __forceinline__ void init() { item = blockIdx.x, tile = chunk = 0, n = tcount = it = 0; }
  __device__ __forceinline__ void advance(const AttnBwdTcParams& p) {
    ++n;
    if (++chunk == p.nc) {
      chunk = 0, ++tcount;
      if (++tile == p.nt) tile = 0, item += gridDim.x, ++it;
    }
  }
  __device__ __forceinline__ bool valid(const AttnBwdTcParams& p) const { return item < p.n_items; }
};

__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// One row x `wc` score columns (wc % 16 == 0, <= 96):  dS (and P) -> packed bf16, written back in
// place.  16-column register blocks, double-buffered: the TMEM loads of block b+1 are in flight while
// block b is computed (tcgen05.wait::ld only covers loads issued before it).
// DKV: per-COLUMN statistics from shared memory (stat_saddr -> -lse*log2e of column c0; delta
// BW_STAT floats further);  DQ: per-row statistics in registers.
template <bool DKV>
__device__ __forceinline__ void bwd_chunk(uint32_t t_s, uint32_t t_dp, uint32_t stat_saddr, float nl_row,
                                          float de_row, int c0, int wc, int N) {
  uint32_t s[2][16], dp[2][16];
  tmem_ld_32x32b_x16(t_s, s[0]);
  tmem_ld_32x32b_x16(t_dp, dp[0]);
  const int nb = wc >> 4;
#pragma unroll
  for (int b = 0; b < BW_CW / 16; ++b) {
    if (b < nb) {   // warp-uniform
      tmem_ld_wait();
      if (b + 1 < nb) {
        tmem_ld_32x32b_x16(t_s + (b + 1) * 16, s[(b + 1) & 1]);
        tmem_ld_32x32b_x16(t_dp + (b + 1) * 16, dp[(b + 1) & 1]);
      }
      uint32_t pp[8], dd[8];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float nl[4], de[4];
        if constexpr (DKV) {
          const float4 x = lds_f4(stat_saddr + (b * 16 + j) * 4);                  // broadcast reads
          const float4 y = lds_f4(stat_saddr + (BW_STAT + b * 16 + j) * 4);
          nl[0] = x.x, nl[1] = x.y, nl[2] = x.z, nl[3] = x.w;
          de[0] = y.x, de[1] = y.y, de[2] = y.z, de[3] = y.w;
        } else {
          nl[0] = nl[1] = nl[2] = nl[3] = nl_row;
          de[0] = de[1] = de[2] = de[3] = de_row;
        }
        float pv[4], dv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pv[e] = fast_ex2(fmaf(__uint_as_float(s[b & 1][j + e]), kLog2eBw, nl[e]));
          dv[e] = pv[e] * (__uint_as_float(dp[b & 1][j + e]) - de[e]);
        }
        if constexpr (!DKV) {
          // key columns past N hold zero-filled K rows; keep 0 * inf out of the dQ accumulation
          if (c0 + b * 16 + 16 > N) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c0 + b * 16 + j + e >= N) dv[e] = 0.f;
          }
        }
        pp[j >> 1] = pack_bf16x2(pv[0], pv[1]), pp[(j >> 1) + 1] = pack_bf16x2(pv[2], pv[3]);
        dd[j >> 1] = pack_bf16x2(dv[0], dv[1]), dd[(j >> 1) + 1] = pack_bf16x2(dv[2], dv[3]);
      }
      if constexpr (DKV) tmem_st_32x32b_x8(t_s + b * 8, pp);
      tmem_st_32x32b_x8(t_dp + b * 8, dd);
    }
  }
}

// accumulator rows (one per thread, 64 fp32 columns in two 32-column register blocks) -> bf16 -> HBM
__device__ __forceinline__ void store_rows64_bf16(uint32_t stage, int lane, const uint32_t (&a)[32], const uint32_t (&b)[32],
                                                  float scale, __nv_bfloat16* g, long ld, int nrows) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(a[j]) * scale;
  warp_store_tile32_bf16(stage, lane, v, g, ld, nrows);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(b[j]) * scale;
  warp_store_tile32_bf16(stage, lane, v, g + 32, ld, nrows);
}

// tm128 / tm16: [3D cols, N rows, n_seq] views of qkv with 64 x 128 and 64 x 16 boxes; td128 / td16:
// the same for d_out (D cols).
template <bool DKV, bool CO>
__global__ void MISSM_PERSISTENT_BOUNDS(CO)   // BW_THREADS threads, one CTA per SM
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16,
                   const __grid_constant__ CUtensorMap td128, const __grid_constant__ CUtensorMap td16,
                   const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // resident column-side operands: [item buffer][operand]
  uint8_t* sRes = smem;                                   // 2 x 2 x 34 KB
  uint8_t* sTile = sRes + 4 * BW_RES_BYTES;               // [tile buffer][operand]: 2 x 2 x 16 KB
  float* sStat = reinterpret_cast<float*>(sTile + 4 * BW_TILE_BYTES);   // [item buffer][nlse2 | delta][BW_STAT]
  uint8_t* sStage = reinterpret_cast<uint8_t*>(sStat + 4 * BW_STAT);    // 8 softmax warps x 2 KB output staging
  AttnBwdSmem* sh = reinterpret_cast<AttnBwdSmem*>(sStage + 8 * 2048);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NACC = DKV ? 1 : 2;   // DKV: dV | dK fill the 128 accumulator columns; DQ: dQ double-buffered

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm128), tma_prefetch_desc(&tm16), tma_prefetch_desc(&td128), tma_prefetch_desc(&td16);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh->res_full[i], 1), mbar_init(&sh->res_empty[i], p.tail ? 3 : 1);   // MMA commit (+ the two tail warps)
      mbar_init(&sh->a_full[i], 1), mbar_init(&sh->a_empty[i], 1);
      mbar_init(&sh->s_full[i], 1), mbar_init(&sh->p_full[i], 128), mbar_init(&sh->sbuf_empty[i], 1);
      mbar_init(&sh->acc_full[i], 1), mbar_init(&sh->acc_empty[i], 128);
      mbar_init(&sh->stat_full[i], 32), mbar_init(&sh->stat_empty[i], 256);
    }
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
  // register pool of the CTA = 384 x kCoResidentRegs: control warpgroup 104, softmax warpgroups 184 (see attention_tc.cu)
  const int n_full = p.sw / 128;             // full 128-row boxes of a resident operand
  const int n_rem16 = (p.sw % 128) / 16;     // remaining 16-row boxes
  // column offsets (elements) of the operands inside a qkv row
  const int col_q = 0, col_k = p.D, col_v = 2 * p.D;

  // register pool of the CTA = 384 x kCoResidentRegs: control warpgroup 104, softmax warpgroups 184 (see attention_tc.cu)
  if (warp < 4) {
  if constexpr (CO) reg_dealloc<104>();
  if (warp == 0) {
    // ================================ TMA producer ====================================
    // (converged warp waits; one elected lane arms the barrier and issues the copies)
    {
      uint32_t it = 0, tcount = 0;
      const uint32_t res_tx = 2u * static_cast<uint32_t>(n_full * BW_TILE_BYTES + n_rem16 * 16 * 128);
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int s = item / p.H, h = item % p.H;
        const int rb = it & 1;
        mbar_wait(&sh->res_empty[rb], ((it >> 1) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&sh->res_full[rb], res_tx);
          uint8_t* r0 = sRes + (rb * 2 + 0) * BW_RES_BYTES;
          uint8_t* r1 = sRes + (rb * 2 + 1) * BW_RES_BYTES;
          // DKV: resident = Q, dO.   DQ: resident = K, V.
          for (int j = 0; j < n_full; ++j) {
            tma_load_3d(r0 + j * BW_TILE_BYTES, &tm128, &sh->res_full[rb], (DKV ? col_q : col_k) + h * 64, j * 128, s);
            if (DKV) tma_load_3d(r1 + j * BW_TILE_BYTES, &td128, &sh->res_full[rb], h * 64, j * 128, s);
            else     tma_load_3d(r1 + j * BW_TILE_BYTES, &tm128, &sh->res_full[rb], col_v + h * 64, j * 128, s);
          }
          for (int j = 0; j < n_rem16; ++j) {
            const int row = n_full * 128 + j * 16;
            tma_load_3d(r0 + row * 128, &tm16, &sh->res_full[rb], (DKV ? col_q : col_k) + h * 64, row, s);
            if (DKV) tma_load_3d(r1 + row * 128, &td16, &sh->res_full[rb], h * 64, row, s);
            else     tma_load_3d(r1 + row * 128, &tm16, &sh->res_full[rb], col_v + h * 64, row, s);
          }
        }
        __syncwarp();
        for (int t = 0; t < p.nt; ++t, ++tcount) {
          const int ab = tcount & 1;
          mbar_wait(&sh->a_empty[ab], ((tcount >> 1) & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&sh->a_full[ab], 2 * BW_TILE_BYTES);
            uint8_t* a0 = sTile + (ab * 2 + 0) * BW_TILE_BYTES;
            uint8_t* a1 = sTile + (ab * 2 + 1) * BW_TILE_BYTES;
            // DKV: tiles = K, V.   DQ: tiles = Q, dO.
            tma_load_3d(a0, &tm128, &sh->a_full[ab], (DKV ? col_k : col_q) + h * 64, t * 128, s);
            if (DKV) tma_load_3d(a1, &tm128, &sh->a_full[ab], col_v + h * 64, t * 128, s);
            else     tma_load_3d(a1, &td128, &sh->a_full[ab], h * 64, t * 128, s);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ======================================
    // the whole (converged) warp walks the schedule and waits; one elected lane issues
    {
      const uint32_t idesc_acc = umma_idesc_bf16_f32(128, 64, 0, 1);   // A from TMEM, B MN-major
      // K-major (score MMAs) and MN-major (accumulation MMAs) views of a tile share one encoding
      const uint64_t desc_tile = umma_smem_desc_sw128(smem_u32(sTile), 16, 1024);
      const uint64_t desc_res = umma_smem_desc_sw128(smem_u32(sRes), 16, 1024);
      uint32_t tr = 0;
      auto issue_scores = [&](const BwCursor& c) {
        const int rb = c.it & 1, ab = c.tcount & 1, sb = c.n & 1;
        if (c.chunk == 0) {
          if (c.tile == 0) mbar_wait(&sh->res_full[rb], (c.it >> 1) & 1);
          mbar_wait(&sh->a_full[ab], (c.tcount >> 1) & 1);
        }
        mbar_wait(&sh->sbuf_empty[sb], ((c.n >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) bw_trace(p, 0, tr, 1, c.n);
        const int c0 = c.chunk * BW_CW;
        const int wc = min(BW_CW, p.sw - c0);
        const uint32_t idesc = umma_idesc_bf16_f32(128, wc, 0, 0);
        // descriptors differ only in the 16-byte-granular start address: base + (byte offset >> 4)
        const uint64_t a0 = desc_tile + ((ab * 2 + 0) * (BW_TILE_BYTES >> 4));
        const uint64_t a1 = desc_tile + ((ab * 2 + 1) * (BW_TILE_BYTES >> 4));
        const uint64_t b0 = desc_res + ((rb * 2 + 0) * (BW_RES_BYTES >> 4) + c0 * 8);
        const uint64_t b1 = desc_res + ((rb * 2 + 1) * (BW_RES_BYTES >> 4) + c0 * 8);
        const uint32_t d_s = tmem + sb * 192, d_dp = d_s + 96;
        if (elect_one_sync()) {
          // the two accumulation chains are interleaved: back-to-back MMAs into the SAME TMEM
          // accumulator serialise on its read-modify-write latency (~100 cycles measured for these
          // small shapes), independent chains overlap
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_ss(d_s, a0 + k * 2, b0 + k * 2, idesc, k > 0);
            umma_f16_ss(d_dp, a1 + k * 2, b1 + k * 2, idesc, k > 0);
          }
          umma_commit(&sh->s_full[sb]);
        }
        __syncwarp();
        if (lane == 0) bw_trace(p, 0, tr, 2, c.n);
      };
      auto issue_accum = [&](const BwCursor& c) {
        const int rb = c.it & 1, ab = c.tcount & 1, sb = c.n & 1;
        const int acc = DKV ? 0 : (c.tcount & 1);
        const uint32_t acc_use = DKV ? c.tcount : (c.tcount >> 1);
        mbar_wait(&sh->p_full[sb], (c.n >> 1) & 1);
        if (lane == 0) bw_trace(p, 0, tr, 3, c.n);
        if (c.chunk == 0) mbar_wait(&sh->acc_empty[acc], (acc_use & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) bw_trace(p, 0, tr, 4, c.n);
        const int c0 = c.chunk * BW_CW;
        const int ksteps = min(BW_CW, p.sw - c0) / 16;
        const uint64_t b0 = desc_res + ((rb * 2 + 0) * (BW_RES_BYTES >> 4) + c0 * 8);
        const uint64_t b1 = desc_res + ((rb * 2 + 1) * (BW_RES_BYTES >> 4) + c0 * 8);
        const uint32_t t_p = tmem + sb * 192, t_ds = t_p + 96;
        if (elect_one_sync()) {
          if constexpr (DKV) {
            const uint32_t d_dv = tmem + BW_ACC_COL, d_dk = d_dv + 64;
#pragma unroll
            for (int k = 0; k < BW_CW / 16; ++k) {
              if (k < ksteps) {
                umma_f16_ts(d_dv, t_p + k * 8, b1 + k * 128, idesc_acc, (c.chunk | k) != 0);    // dV += P^T dO
                umma_f16_ts(d_dk, t_ds + k * 8, b0 + k * 128, idesc_acc, (c.chunk | k) != 0);   // dK += dS^T Q
              }
            }
          } else {
            const uint32_t d_dq = tmem + BW_ACC_COL + acc * 64;
#pragma unroll
            for (int k = 0; k < BW_CW / 16; ++k)     // dQ += dS K
              if (k < ksteps) umma_f16_ts(d_dq, t_ds + k * 8, b0 + k * 128, idesc_acc, (c.chunk | k) != 0);
          }
          umma_commit(&sh->sbuf_empty[sb]);
          if (c.chunk == p.nc - 1) {
            umma_commit(&sh->acc_full[acc]);
            umma_commit(&sh->a_empty[ab]);
            if (c.tile == p.nt - 1) umma_commit(&sh->res_empty[rb]);
          }
        }
        __syncwarp();
        if (lane == 0) bw_trace(p, 0, tr, 5, c.n);
      };
      BwCursor cs, cp;
      cs.init(), cp.init();
      if (cs.valid(p)) {
        issue_scores(cs);
        cs.advance(p);
      }
      while (cp.valid(p)) {
        if (cs.valid(p)) {
          issue_scores(cs);
          cs.advance(p);
        }
        issue_accum(cp);
        cp.advance(p);
      }
    }
  } else {
    // ===== warp 3: column statistics (DKV): -lse*log2e and delta per query;  warps 2 + 3: the odd row =====
    if constexpr (DKV) {
      auto load_stats = [&](uint32_t it, int item) {
        const int sbuf = it & 1;
        mbar_wait(&sh->stat_empty[sbuf], ((it >> 1) & 1) ^ 1);
        float* nl = sStat + (sbuf * 2 + 0) * BW_STAT;
        float* de = sStat + (sbuf * 2 + 1) * BW_STAT;
        const long base = static_cast<long>(item) * p.N;     // item = s * H + h
        for (int q = lane; q < BW_STAT; q += 32) {
          const bool ok = q < p.N;
          nl[q] = ok ? -p.lse[base + q] * kLog2eBw : -INFINITY;
          de[q] = ok ? p.delta[base + q] : 0.f;
        }
        mbar_arrive(&sh->stat_full[sbuf]);
      };
      // (the odd KEY row of DKV is not computed here: attn_delta_tail_kernel does it, see below)
      if (warp == 3) {
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) load_stats(it, item);
      }
    } else {
      if (p.tail) {
        // ---- query N-1: dQ[N-1] = q_scale * sum_k dS[k] K[k]
        const int tw = warp - 2, t = tw * 32 + lane;
        const uint32_t w0_s = smem_u32(sh->tail_w[0]), w1_s = smem_u32(sh->tail_w[1]);
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
          const int s = item / p.H, h = item % p.H, rb = it & 1;
          const long row = static_cast<long>(s) * p.N + (p.N - 1);
          const long li = static_cast<long>(item) * p.N + (p.N - 1);
          const float nl = -p.lse[li] * kLog2eBw, de = p.delta[li];
          float a[64], sc[kTailSlots];
          tail_load_row64(p.qkv + row * p.ld_qkv + h * 64, a);             // q (pre-scaled)
          mbar_wait(&sh->res_full[rb], (it >> 1) & 1);
          const uint32_t k_s = smem_u32(sRes + (rb * 2 + 0) * BW_RES_BYTES);
          const uint32_t v_s = smem_u32(sRes + (rb * 2 + 1) * BW_RES_BYTES);
#pragma unroll
          for (int j = 0; j < kTailSlots; ++j) {
            const int r = t + 64 * j;
            sc[j] = r < p.N ? tail_dot64(k_s, r, a) : 0.f;
          }
          tail_load_row64(p.d_out + row * p.ld_o + h * 64, a);             // dO
#pragma unroll
          for (int j = 0; j < kTailSlots; ++j) {
            const int r = t + 64 * j;
            float ds = 0.f;
            if (r < p.N) ds = fast_ex2(fmaf(sc[j], kLog2eBw, nl)) * (tail_dot64(v_s, r, a) - de);
            sts_f32(w0_s + r * 4, ds);
          }
          tail_team_sync();
          float dq[8];
          tail_weighted_rowsum(k_s, w0_s, tw == 0 ? 0 : 128, tw == 0 ? 128 : p.N, lane, dq);
          __syncwarp();                                    // every lane is done with K, V
          if (lane == 0) mbar_arrive(&sh->res_empty[rb]);
          if (tw == 1 && lane < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sts_f32(w1_s + (lane * 8 + j) * 4, dq[j]);
          }
          tail_team_sync();
          if (tw == 0) {
            if (lane < 8) {
#pragma unroll
              for (int j = 0; j < 8; ++j) dq[j] += lds_f32(w1_s + (lane * 8 + j) * 4);
            }
            tail_store_row64(p.dqkv + row * p.ld_qkv + h * 64, lane, dq, p.q_scale);
          }
        }
      }
    }
  }
  } else {
    if constexpr (CO) reg_alloc<184>();
    // ========================= softmax warpgroups (one thread per row) =================
    const int g = (warp - 4) >> 2;       // warpgroup: takes the steps with n % 2 == g
    const int q = warp & 3;              // TMEM lane quarter
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float nl_row = 0.f, de_row = 0.f, nl_pre = 0.f, de_pre = 0.f;
    int stats_tile = -1, pre_tile = -1;
    uint32_t tr = 0;
    const bool tracer = (q == 0 && lane == 0);
    const uint32_t stage = smem_u32(sStage) + (warp - 4) * 2048;
    BwCursor c;
    c.init();
    while (c.valid(p)) {
      const uint32_t it_now = c.it;
      if constexpr (DKV) {
        // EVERY thread waits for the statistics of an item before it may release them, also the warpgroup that
        // has no step in this item (one step per item when N <= 96): otherwise that warpgroup could arrive on
        // stat_empty for item it+2 before the other one has arrived for item it, the barrier phase would
        // complete early, the loader would overwrite statistics still in use and the other warpgroup could
        // miss a whole phase of stat_full (parity aliasing -> deadlock)
        if (c.tile == 0 && c.chunk == 0) mbar_wait(&sh->stat_full[c.it & 1], (c.it >> 1) & 1);
      }
      if ((c.n & 1) == static_cast<uint32_t>(g)) {
        const int row = c.tile * 128 + q * 32 + lane;
        const bool warp_has_rows = c.tile * 128 + q * 32 < p.N;
        uint32_t stat_saddr = 0;
        if constexpr (DKV) {
          stat_saddr = smem_u32(sStat + (c.it & 1) * 2 * BW_STAT);
        } else {
          if (stats_tile != static_cast<int>(c.tcount)) {
            // this tile's row statistics were requested one tile ago; request the next tile's now
            stats_tile = static_cast<int>(c.tcount);
            if (pre_tile == stats_tile) {
              nl_row = nl_pre, de_row = de_pre;
            } else {
              const bool ok = row < p.N;
              const long li = static_cast<long>(c.item) * p.N + (ok ? row : 0);
              nl_row = ok ? p.lse[li] : INFINITY;
              de_row = ok ? p.delta[li] : 0.f;
            }
            nl_row *= -kLog2eBw;
            const int nx_tile = (c.tile + 1 < p.nt) ? c.tile + 1 : 0;
            const int nx_item = (c.tile + 1 < p.nt) ? c.item : c.item + static_cast<int>(gridDim.x);
            pre_tile = stats_tile + 1;
            const int nx_row = nx_tile * 128 + q * 32 + lane;
            const bool ok = nx_row < p.N && nx_item < p.n_items;
            const long li = ok ? static_cast<long>(nx_item) * p.N + nx_row : 0;
            nl_pre = ok ? p.lse[li] : INFINITY;
            de_pre = ok ? p.delta[li] : 0.f;
          }
        }
        if (tracer) bw_trace(p, 1 + g, tr, 10, c.n);
        mbar_wait(&sh->s_full[g], (c.n >> 1) & 1);
        tc_fence_after();
        if (tracer) bw_trace(p, 1 + g, tr, 11, c.n);
        const int c0 = c.chunk * BW_CW;
        const int wc = min(BW_CW, p.sw - c0);
        if (warp_has_rows) {
          const uint32_t t_s = tmem + lane_addr + g * 192, t_dp = t_s + 96;
          bwd_chunk<DKV>(t_s, t_dp, stat_saddr + c0 * 4, nl_row, de_row, c0, wc, p.N);
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&sh->p_full[g]);
        if (tracer) bw_trace(p, 1 + g, tr, 12, c.n);

        if (c.chunk == p.nc - 1) {
          // ---- accumulators of this row tile -> HBM (this warpgroup handled the tile's last chunk)
          const int acc = DKV ? 0 : (c.tcount & 1);
          const uint32_t acc_use = DKV ? c.tcount : (c.tcount >> 1);
          mbar_wait(&sh->acc_full[acc], acc_use & 1);
          tc_fence_after();
          if (tracer) bw_trace(p, 1 + g, tr, 13, c.n);
          const int s = c.item / p.H, h = c.item % p.H;
          const int wrow0 = c.tile * 128 + q * 32;            // first row of this warp
          __nv_bfloat16* grow = p.dqkv + (static_cast<long>(s) * p.N + wrow0) * p.ld_qkv + h * 64;
          const int nrows = p.N - wrow0;
          if constexpr (DKV) {
            uint32_t x0[32], x1[32];
            if (warp_has_rows) {
              tmem_ld_32x32b_x32(tmem + lane_addr + BW_ACC_COL, x0);
              tmem_ld_32x32b_x32(tmem + lane_addr + BW_ACC_COL + 32, x1);
              tmem_ld_wait();
              store_rows64_bf16(stage, lane, x0, x1, 1.0f, grow + 2 * p.D, p.ld_qkv, nrows);     // dV
              tmem_ld_32x32b_x32(tmem + lane_addr + BW_ACC_COL + 64, x0);
              tmem_ld_32x32b_x32(tmem + lane_addr + BW_ACC_COL + 96, x1);
              tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(&sh->acc_empty[acc]);
            if (warp_has_rows) store_rows64_bf16(stage, lane, x0, x1, 1.0f, grow + p.D, p.ld_qkv, nrows);   // dK
          } else {
            uint32_t x0[32], x1[32];
            if (warp_has_rows) {
              tmem_ld_32x32b_x32(tmem + lane_addr + BW_ACC_COL + acc * 64, x0);
              tmem_ld_32x32b_x32(tmem + lane_addr + BW_ACC_COL + acc * 64 + 32, x1);
              tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(&sh->acc_empty[acc]);
            if (warp_has_rows) store_rows64_bf16(stage, lane, x0, x1, p.q_scale, grow, p.ld_qkv, nrows);    // dQ
          }
          if (tracer) bw_trace(p, 1 + g, tr, 14, c.n);
        }
      }
      c.advance(p);
      if constexpr (DKV) {
        if (c.it != it_now) mbar_arrive(&sh->stat_empty[it_now & 1]);   // done with that item's statistics
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------
// Backward preprocess of the tcgen05 path, one CTA per (sequence, head):
//   delta[q] = sum_d dO[q, d] * O[q, d]                          (all rows)
// and, when the odd row N-1 is kept out of the DKV tiles (tail = 1), everything that belongs to KEY N-1:
//   p[q] = exp(q . k[N-1] - lse[q]),  dS[q] = p[q] (dO[q] . v[N-1] - delta[q])
//   dV[N-1] = sum_q p[q] dO[q],       dK[N-1] = sum_q dS[q] Q[q]
// It needs nothing the main kernels produce, so it rides on the pass that reads dO anyway, spread over all
// SMs (inside the DKV kernel the same work on two spare warps cost more than the third tile it replaced).
// 8 lanes per row: a row of one head is 128 contiguous bytes -> one 16-byte load per lane and operand.
// HBM/L2-bound: reads O, dO (+ Q with tail) once.
// ---------------------------------------------------------------------------------------
struct AttnDeltaTailParams {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* out;
  const __nv_bfloat16* d_out;
  const float* lse;
  float* delta;
  __nv_bfloat16* dqkv;
  long ld_qkv, ld_o;
  int N, H, D, tail;
};

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y, v[4] = c.x, v[5] = c.y, v[6] = d.x, v[7] = d.y;
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
  return ((a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3])) + ((a[4] * b[4] + a[5] * b[5]) + (a[6] * b[6] + a[7] * b[7]));
}
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

constexpr int DT_THREADS = 256;   // 32 row groups of 8 lanes

__global__ void __launch_bounds__(DT_THREADS)
attn_delta_tail_kernel(const AttnDeltaTailParams p) {
  __shared__ float red[2][DT_THREADS / 8][64];
  const int item = blockIdx.x, s = item / p.H, h = item % p.H;
  const int grp = threadIdx.x >> 3, gl = threadIdx.x & 7;
  const long row0 = static_cast<long>(s) * p.N;
  const __nv_bfloat16* o_base = p.out + row0 * p.ld_o + h * 64 + gl * 8;
  const __nv_bfloat16* do_base = p.d_out + row0 * p.ld_o + h * 64 + gl * 8;
  const __nv_bfloat16* q_base = p.qkv + row0 * p.ld_qkv + h * 64 + gl * 8;
  float kt[8], vt[8], acc_v[8], acc_k[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) kt[j] = vt[j] = acc_v[j] = acc_k[j] = 0.f;
  if (p.tail) {
    const __nv_bfloat16* t_row = q_base + static_cast<long>(p.N - 1) * p.ld_qkv;
    unpack8(__ldg(reinterpret_cast<const uint4*>(t_row + p.D)), kt);
    unpack8(__ldg(reinterpret_cast<const uint4*>(t_row + 2 * p.D)), vt);
  }
  for (int q0 = 0; q0 < p.N; q0 += DT_THREADS / 8) {     // warp-uniform trip count (shuffles inside)
    const int q = q0 + grp;
    const bool valid = q < p.N;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    float o8[8], d8[8];
    unpack8(valid ? __ldg(reinterpret_cast<const uint4*>(o_base + q * p.ld_o)) : zero, o8);
    unpack8(valid ? __ldg(reinterpret_cast<const uint4*>(do_base + q * p.ld_o)) : zero, d8);
    uint4 qraw = zero;
    if (p.tail && valid) qraw = __ldg(reinterpret_cast<const uint4*>(q_base + q * p.ld_qkv));
    const float de = group8_sum(dot8(o8, d8));
    const long li = static_cast<long>(item) * p.N + (valid ? q : 0);
    if (gl == 0 && valid) p.delta[li] = de;
    if (p.tail) {
      float q8[8];
      unpack8(qraw, q8);
      const float sd = group8_sum(dot8(kt, q8));
      const float dp = group8_sum(dot8(vt, d8));
      const float pr = valid ? fast_ex2((sd - p.lse[li]) * kLog2eBw) : 0.f;
      const float ds = pr * (dp - de);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc_v[j] = fmaf(pr, d8[j], acc_v[j]), acc_k[j] = fmaf(ds, q8[j], acc_k[j]);
    }
  }
  if (p.tail) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[0][grp][gl * 8 + j] = acc_v[j], red[1][grp][gl * 8 + j] = acc_k[j];
    __syncthreads();
    if (threadIdx.x < 128) {
      const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
      float t = 0.f;
#pragma unroll 8
      for (int g = 0; g < DT_THREADS / 8; ++g) t += red[which][g][c];
      __nv_bfloat16* dst = p.dqkv + (row0 + p.N - 1) * p.ld_qkv + (which == 0 ? 2 * p.D : p.D) + h * 64 + c;
      *dst = __float2bfloat16(t);
    }
  }
}

// returns 0 if launched, -1 if the shape is not handled here (caller uses the general mma.sync
// path), > 0 on error.  Fills delta itself (attn_delta_tail_kernel).
int attention_bwd_tc(const missm_attn_args* a, cudaStream_t stream) {
  // N <= 96 (one column chunk = one step per item) is declined: the two softmax warpgroups then alternate over
  // whole ITEMS, and a warpgroup that waits on acc_full only every second phase can see the phase before last
  // as "its" parity (mbarrier parity aliasing) -- with >= 2 steps per tile the in-order MMA commits exclude that
  const bool ok = !a->causal && a->key_mask == nullptr && a->s_in == 1 && a->tok_stride == 1 &&
                  a->seq_outer == a->N && a->N <= BW_MAXN && a->N > BW_CW && a->head_dim == 64;
  if (!ok) return -1;
  CUtensorMap tm128, tm16, td128, td16;
  const uint64_t seq_q = static_cast<uint64_t>(a->N) * a->ld_qkv, seq_o = static_cast<uint64_t>(a->N) * a->ld_o;
  if (int rc = make_tmap_3d_bf16(&tm128, a->qkv, 3 * static_cast<uint64_t>(a->D), a->N, a->n_seq, a->ld_qkv, seq_q, 64, 128)) return rc;
  if (int rc = make_tmap_3d_bf16(&tm16, a->qkv, 3 * static_cast<uint64_t>(a->D), a->N, a->n_seq, a->ld_qkv, seq_q, 64, 16)) return rc;
  if (int rc = make_tmap_3d_bf16(&td128, a->d_out, a->D, a->N, a->n_seq, a->ld_o, seq_o, 64, 128)) return rc;
  if (int rc = make_tmap_3d_bf16(&td16, a->d_out, a->D, a->N, a->n_seq, a->ld_o, seq_o, 64, 16)) return rc;
  AttnBwdTcParams p;
  p.N = a->N, p.H = a->H, p.D = a->D, p.n_items = a->n_seq * a->H;
  p.sw = (a->N + 15) / 16 * 16;
  p.tail = attention_tail_enabled(a->N) ? 1 : 0;
  p.nt = p.tail ? a->N / 128 : (a->N + 127) / 128;
  p.nc = (p.sw + BW_CW - 1) / BW_CW;
  p.qkv = static_cast<const __nv_bfloat16*>(a->qkv), p.d_out = static_cast<const __nv_bfloat16*>(a->d_out), p.ld_o = a->ld_o;
  p.lse = a->lse, p.delta = a->delta;
  p.dqkv = static_cast<__nv_bfloat16*>(a->dqkv), p.ld_qkv = a->ld_qkv, p.q_scale = a->q_scale;
  const int smem = 4 * BW_RES_BYTES + 4 * BW_TILE_BYTES + 4 * BW_STAT * 4 + 8 * 2048 + static_cast<int>(sizeof(AttnBwdSmem)) + 1024;
  static bool configured = false;
  if (!configured) {
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int grid = p.n_items < persistent_sms() ? p.n_items : persistent_sms();
  auto launch_delta = [&](int tail) {
    AttnDeltaTailParams d;
    d.qkv = p.qkv, d.out = static_cast<const __nv_bfloat16*>(a->out), d.d_out = p.d_out, d.lse = a->lse;
    d.delta = a->delta, d.dqkv = p.dqkv, d.ld_qkv = a->ld_qkv, d.ld_o = a->ld_o;
    d.N = a->N, d.H = a->H, d.D = a->D, d.tail = tail;
    attn_delta_tail_kernel<<<p.n_items, DT_THREADS, 0, stream>>>(d); note_launch();
  };
  p.trace = nullptr;
  const char* trace_path = getenv("MISSM_ATTN_TRACE");   // debugging aid: dumps CTA 0's event clocks (synchronises!)
  if (trace_path != nullptr) {
    const size_t nb = 2 * 3 * 330 * 3 * sizeof(long long);
    long long* d = nullptr;
    MISSM_CHECK_CUDA(cudaMalloc(&d, nb));
    MISSM_CHECK_CUDA(cudaMemsetAsync(d, 0, nb, stream));
    p.trace = d;
    launch_delta(0);
    p.tail = 0, p.nt = (a->N + 127) / 128;     // tracing walks all tiles
    attn_bwd_tc_kernel<true, false><<<grid, BW_THREADS, smem, stream>>>(tm128, tm16, td128, td16, p); note_launch();
    p.trace = d + 3 * 330 * 3;
    attn_bwd_tc_kernel<false, false><<<grid, BW_THREADS, smem, stream>>>(tm128, tm16, td128, td16, p); note_launch();
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
  launch_delta(p.tail);
  // DKV: its odd KEY row comes from attn_delta_tail_kernel (the kernel only walks the full tiles);
  // DQ: its odd QUERY row is computed by warps 2-3 of the kernel itself
  AttnBwdTcParams pk = p;
  pk.tail = 0;
  if (coresident()) {
    attn_bwd_tc_kernel<true, true><<<grid, BW_THREADS, smem, stream>>>(tm128, tm16, td128, td16, pk); note_launch();
    attn_bwd_tc_kernel<false, true><<<grid, BW_THREADS, smem, stream>>>(tm128, tm16, td128, td16, p); note_launch();
  } else {
    attn_bwd_tc_kernel<true, false><<<grid, BW_THREADS, smem, stream>>>(tm128, tm16, td128, td16, pk); note_launch();
    attn_bwd_tc_kernel<false, false><<<grid, BW_THREADS, smem, stream>>>(tm128, tm16, td128, td16, p); note_launch();
  }
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace missm
