// Fused attention FORWARD on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), head_dim 64,
// non-causal, unmasked, whole key range resident (N <= 272): the ViT-L/14 spatial attention of the
// image / depth / thermal / video towers (N = 257), i.e. transformers 4.3x CLIPAttention's
// bmm -> softmax -> bmm chain called at languagebind/image/modeling_image.py:140.
//
// One persistent CTA per SM walks (sequence, head) items.  Per item K and V (all N keys, bf16) are
// TMA-loaded once into 128B-swizzled shared memory; per 128-row query tile:
//   warp 1 : S[128 x Nk] = Q K^T        tcgen05.mma, A/B from smem (K-major), fp32 in TMEM
//   warps 4-7 (one thread per query row): row max + exp2 + row sum straight from TMEM
//            (tcgen05.ld), P written back IN PLACE as packed bf16 (tcgen05.st) -- S never leaves
//            the SM and P never touches shared memory
//   warp 1 : O[128 x 64] = P V          tcgen05.mma with the A operand read from TMEM, V as an
//            MN-major smem operand (no transpose of V anywhere)
//   warps 4-7: O / rowsum -> bf16 -> HBM (one 128 B line per row), lse -> HBM
// Because the whole key range fits in TMEM (272 of 512 columns) the softmax is exact two-pass, no
// online rescaling.  The S MMA of tile t+1 overlaps the epilogue of tile t.
//
// Bound: MUFU (one ex2 per score) / tensor.  Algorithmic work per item = 4 * N^2 * 64 flop.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

constexpr int kTcThreads = 256;
constexpr int TC_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int TC_MAX_KV = 272;            // keys (multiple of 16) that fit the single-pass design
constexpr int TC_S_COL = 0;               // TMEM columns: S / P at [0, 272), O at [384, 448)
constexpr int TC_O_COL = 384;
constexpr float kLog2eTc = 1.4426950408889634f;

struct AttnTcParams {
  int N, H, D, n_items;   // tokens per sequence, heads, model width, n_seq * H
  int sw;                 // N rounded up to 16
  __nv_bfloat16* out;
  long ld_o;
  float* lse;             // [n_seq, H, N]
};

struct AttnTcSmem {
  uint64_t kv_full, kv_empty;
  uint64_t q_full[2], q_empty[2];
  uint64_t s_full, p_full, o_full, o_empty;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                          // 2 x 16 KB
  uint8_t* sK = smem + 2 * TC_TILE_BYTES;      // 3 x 16 KB (rows 0..383, zero past N)
  uint8_t* sV = sK + 3 * TC_TILE_BYTES;        // 3 x 16 KB
  AttnTcSmem* sh = reinterpret_cast<AttnTcSmem*>(sV + 3 * TC_TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (p.N + 127) / 128;            // query tiles per item

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(&sh->kv_full, 1), mbar_init(&sh->kv_empty, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&sh->q_full[i], 1), mbar_init(&sh->q_empty[i], 1);
    mbar_init(&sh->s_full, 1), mbar_init(&sh->o_full, 1);
    mbar_init(&sh->p_full, 128), mbar_init(&sh->o_empty, 128);
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

  if (warp == 0) {
    // ================================ TMA producer ====================================
    if (lane == 0) {
      uint32_t it = 0, qcount = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int s = item / p.H, h = item % p.H;
        mbar_wait(&sh->kv_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(&sh->kv_full, 6 * TC_TILE_BYTES);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          tma_load_3d(sK + j * TC_TILE_BYTES, &tm_qkv, &sh->kv_full, p.D + h * 64, j * 128, s);
          tma_load_3d(sV + j * TC_TILE_BYTES, &tm_qkv, &sh->kv_full, 2 * p.D + h * 64, j * 128, s);
        }
        for (int t = 0; t < nt; ++t, ++qcount) {
          const int buf = qcount & 1;
          mbar_wait(&sh->q_empty[buf], ((qcount >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&sh->q_full[buf], TC_TILE_BYTES);
          tma_load_3d(sQ + buf * TC_TILE_BYTES, &tm_qkv, &sh->q_full[buf], h * 64, t * 128, s);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ======================================
    if (lane == 0) {
      const int n1 = p.sw > 256 ? 256 : p.sw;   // first S MMA width
      const int n2 = p.sw - n1;                 // second (0 or 16)
      const uint32_t idesc_s1 = umma_idesc_bf16_f32(128, n1, 0, 0);
      const uint32_t idesc_s2 = umma_idesc_bf16_f32(128, n2 > 0 ? n2 : 16, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16_f32(128, 64, 0, 1);   // A from TMEM, B = V MN-major
      const uint32_t k_base = smem_u32(sK), v_base = smem_u32(sV);
      uint32_t it = 0, qcount = 0, tcount = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        mbar_wait(&sh->kv_full, it & 1);
        for (int t = 0; t < nt; ++t, ++qcount, ++tcount) {
          const int buf = qcount & 1;
          mbar_wait(&sh->q_full[buf], (qcount >> 1) & 1);
          tc_fence_after();
          const uint32_t q_base = smem_u32(sQ + buf * TC_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a = umma_smem_desc_sw128(q_base + k * 32, 16, 1024);
            umma_f16_ss(tmem + TC_S_COL, a, umma_smem_desc_sw128(k_base + k * 32, 16, 1024), idesc_s1, k > 0);
            if (n2 > 0)
              umma_f16_ss(tmem + TC_S_COL + 256, a,
                          umma_smem_desc_sw128(k_base + 256 * 128 + k * 32, 16, 1024), idesc_s2, k > 0);
          }
          umma_commit(&sh->s_full);
          umma_commit(&sh->q_empty[buf]);
          mbar_wait(&sh->p_full, tcount & 1);          // P is in TMEM
          mbar_wait(&sh->o_empty, (tcount & 1) ^ 1);   // previous O has been read out
          tc_fence_after();
          const int ksteps = p.sw / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_f16_ts(tmem + TC_O_COL, tmem + TC_S_COL + k * 8,
                        umma_smem_desc_sw128(v_base + k * 2048, 16, 1024), idesc_pv, k > 0);
          umma_commit(&sh->o_full);
          if (t == nt - 1) umma_commit(&sh->kv_empty);
        }
      }
    }
  } else if (warp >= 4) {
    // ========================= softmax + output (one thread per query row) =============
    const int q = warp & 3;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int ngroups = p.sw / 16;              // 16-column groups of S
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int s = item / p.H, h = item % p.H;
      for (int t = 0; t < nt; ++t, ++tcount) {
        const int row = t * 128 + q * 32 + lane;
        mbar_wait(&sh->s_full, tcount & 1);
        tc_fence_after();
        const uint32_t s_addr = tmem + lane_addr + TC_S_COL;
        // ---- pass 1: row maximum over the valid keys
        float mx = -INFINITY;
        for (int g = 0; g + 1 < ngroups; g += 2) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(s_addr + g * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (g * 16 + j < p.N) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        if (ngroups & 1) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(s_addr + (ngroups - 1) * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if ((ngroups - 1) * 16 + j < p.N) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        const float mx2 = mx * kLog2eTc;
        // ---- pass 2: P = exp2(S*log2e - max*log2e) -> bf16, in place; row sum
        float sum = 0.f;
        for (int g = 0; g + 1 < ngroups; g += 2) {
          uint32_t r[32], pk[16];
          tmem_ld_32x32b_x32(s_addr + g * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float a = (g * 16 + j < p.N) ? exp2f(fmaf(__uint_as_float(r[j]), kLog2eTc, -mx2)) : 0.f;
            float b = (g * 16 + j + 1 < p.N) ? exp2f(fmaf(__uint_as_float(r[j + 1]), kLog2eTc, -mx2)) : 0.f;
            sum += a + b;
            pk[j >> 1] = pack_bf16x2(a, b);
          }
          tmem_st_32x32b_x16(s_addr + g * 8, pk);
        }
        if (ngroups & 1) {
          const int g = ngroups - 1;
          uint32_t r[16], pk[8];
          tmem_ld_32x32b_x16(s_addr + g * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float a = (g * 16 + j < p.N) ? exp2f(fmaf(__uint_as_float(r[j]), kLog2eTc, -mx2)) : 0.f;
            float b = (g * 16 + j + 1 < p.N) ? exp2f(fmaf(__uint_as_float(r[j + 1]), kLog2eTc, -mx2)) : 0.f;
            sum += a + b;
            pk[j >> 1] = pack_bf16x2(a, b);
          }
          tmem_st_32x32b_x8(s_addr + g * 8, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&sh->p_full);
        // ---- output tile
        mbar_wait(&sh->o_full, tcount & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(tmem + lane_addr + TC_O_COL, o0);
        tmem_ld_32x32b_x32(tmem + lane_addr + TC_O_COL + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sh->o_empty);
        if (row < p.N) {
          const float inv = 1.0f / sum;
          const long grow = static_cast<long>(s) * p.N + row;
          uint4* dst = reinterpret_cast<uint4*>(p.out + grow * p.ld_o + h * 64);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o0[j]) * inv, __uint_as_float(o0[j + 1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(o0[j + 2]) * inv, __uint_as_float(o0[j + 3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(o0[j + 4]) * inv, __uint_as_float(o0[j + 5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(o0[j + 6]) * inv, __uint_as_float(o0[j + 7]) * inv);
            dst[j >> 3] = v;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o1[j]) * inv, __uint_as_float(o1[j + 1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(o1[j + 2]) * inv, __uint_as_float(o1[j + 3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(o1[j + 4]) * inv, __uint_as_float(o1[j + 5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(o1[j + 6]) * inv, __uint_as_float(o1[j + 7]) * inv);
            dst[4 + (j >> 3)] = v;
          }
          if (p.lse != nullptr) p.lse[(static_cast<long>(s) * p.H + h) * p.N + row] = mx + __logf(sum);
        }
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

// returns 0 if launched, -1 if the shape is not handled by this kernel (caller falls back to the
// general mma.sync path), > 0 on error
int attention_fwd_tc(const missm_attn_args* a, cudaStream_t stream) {
  const bool ok = !a->causal && a->key_mask == nullptr && a->s_in == 1 && a->tok_stride == 1 &&
                  a->seq_outer == a->N && a->N <= TC_MAX_KV && a->N >= 16 && a->head_dim == 64;
  if (!ok) return -1;
  CUtensorMap tm;
  if (int rc = make_tmap_3d_bf16(&tm, a->qkv, 3 * static_cast<uint64_t>(a->D), a->N, a->n_seq, a->ld_qkv,
                                 static_cast<uint64_t>(a->N) * a->ld_qkv, 64, 128))
    return rc;
  AttnTcParams p;
  p.N = a->N, p.H = a->H, p.D = a->D, p.n_items = a->n_seq * a->H;
  p.sw = (a->N + 15) / 16 * 16;
  p.out = static_cast<__nv_bfloat16*>(a->out), p.ld_o = a->ld_o, p.lse = a->lse;
  const int smem = 8 * TC_TILE_BYTES + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
  attn_fwd_tc_kernel<<<grid, kTcThreads, smem, stream>>>(tm, p);
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace missm
