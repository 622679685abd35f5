// CTA-pair tcgen05 GEMM for sm_100a (tcgen05.mma.cta_group::2): the large-M path of missm_gemm_bf16.
//
//   C[m,n] = epilogue( sum_k A[m,k] * B[n,k] ),  bf16 operands, fp32 accumulation in TMEM.
//
// Why a second kernel: with one CTA per tile (gemm_tcgen05.cu) every 128x256x16 MMA reads 12 KB of
// shared memory per 128 tensor cycles while TMA writes the same amount -- 187 B/cycle against a
// 128 B/cycle shared-memory port, which caps that kernel near 2/3 of the tensor peak (measured
// 1.32 PFLOP/s at 8192^3, cuBLAS 1.66).  Here two CTAs on the two SMs of a TPC (cluster of 2) share
// one 256 x BN tile: each loads its own 128 rows of A and only HALF of the B tile, and a single
// MMA issued by the leader drives both tensor cores (M = 256), so shared-memory traffic per SM
// drops by a third and the B operand is fetched from L2 once per pair.
//
// Roles per CTA (384 threads, persistent over pair tiles):
//   warp 0 : TMA producer (both CTAs; transaction bytes complete on the LEADER's full barrier)
//   warp 1 : MMA issuer (leader CTA only; commits multicast to both CTAs' barriers)
//   warp 2 : TMEM allocator (cta_group::2, both CTAs)
//   warps 4..11 : epilogue, each CTA drains its own 128 accumulator rows (same fused epilogues as the
//                 1-CTA kernel); "accumulator free" arrives on the leader's barrier across the cluster
// Tile schedule: DYNAMIC through Blackwell's cluster launch control (CLC).  The grid holds one cluster per work
// unit; the clusters that are resident (74 when the GPU is idle) process their own unit and then keep cancelling
// clusters that have not started yet (clusterlaunchcontrol.try_cancel, response multicast into both CTAs'
// shared memory, ~320 cycles) and take over their units.  A pair that starts late -- because another tower's
// kernel, or an NCCL all-reduce CTA, still holds its SMs -- simply processes fewer tiles instead of stretching
// the kernel by a whole tile time as the static "tile = pair + i * pairs" walk does (kept for stream-K and for
// capped grids, MISSM_GEMM_STATIC=1 forces it).
// Bound: tensor pipe.  Algorithmic work per launch = 2*M*N*K flop.
#include <cstdlib>

#define MISSM_KERNEL_TAG "gemm_2cta"
#include "../../include/missm_b200.h"
#include "gemm_common.cuh"
#include "missm_common.cuh"

namespace missm {

template <int BN>
struct Gemm2Cfg {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 8;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers, tmem slot, tile feed*/ +
                                    kEpiWarps * kEpiStageBytes /*epilogue staging*/;
};

// ---------------------------------------------------------------------------------------
// The tile feed of one CTA: the producer warp obtains work units (statically or from CLC) and publishes them to
// the other roles of ITS CTA through a 4-deep ring + mbarriers; -1 ends the kernel.
// ---------------------------------------------------------------------------------------
struct TileFeedSmem {
  uint4 clc_resp;              // CLC response (written by hardware into both CTAs of the pair)
  uint64_t clc_bar;            // ... completes 16 transaction bytes on this barrier, in every CTA
  uint64_t peer_armed;         // leader's copy: the peer has armed clc_bar and is done with the previous response
  uint64_t full[4], empty[4];
  int ring[4];
};

__device__ __forceinline__ void clc_try_cancel(const uint4* resp, const uint64_t* bar) {
  asm volatile(
      "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
      ::"r"(smem_u32(resp)), "r"(smem_u32(bar)) : "memory");
}
// -> first CTA id (x) of the cancelled cluster, or -1 if nothing was left to cancel
__device__ __forceinline__ int clc_decode(const uint4* resp) {
  uint32_t valid = 0, x = 0, y, z;
  asm volatile(
      "{\n.reg .pred p1;\n.reg .b128 r;\nld.shared.b128 r, [%4];\n"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\nselp.u32 %3, 1, 0, p1;\n"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n}\n"
      : "=r"(x), "=r"(y), "=r"(z), "=r"(valid) : "r"(smem_u32(resp)) : "memory");
  (void)y, (void)z;
  return valid ? static_cast<int>(x) : -1;
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// consumer side of the ring (whole warp calls; one lane releases the slot)
__device__ __forceinline__ int feed_take(TileFeedSmem* f, uint32_t& it) {
  const int slot = it & 3;
  mbar_wait(&f->full[slot], (it >> 2) & 1);
  const int w = *reinterpret_cast<volatile int*>(&f->ring[slot]);
  __syncwarp();
  if (lane_id() == 0) mbar_arrive(&f->empty[slot]);
  ++it;
  return w;
}
// linear work unit -> tile (n fastest) and k-block range
__device__ __forceinline__ void unit_decode(const GemmParams& p, int w, int& tile, int& kb0, int& kb1) {
  const int tiles_mn = p.num_m_blk * p.num_n_blk;
  tile = w % tiles_mn;
  kb0 = (w / tiles_mn) * p.kblk_per_split;
  kb1 = kb0 + p.kblk_per_split < p.num_kblk ? kb0 + p.kblk_per_split : p.num_kblk;
}

template <int BN, int EPI, bool OUT_F32, bool CLC, bool CO>
__global__ void __cluster_dims__(2, 1, 1) MISSM_PERSISTENT_BOUNDS(CO)   // kGemmThreads threads, one CTA per SM
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int B_HALF_BYTES = Cfg::B_HALF_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_HALF_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  TileFeedSmem* feed = reinterpret_cast<TileFeedSmem*>(reinterpret_cast<uint8_t*>(full_bar) + 256);   // 16-byte aligned
  uint8_t* sStage = reinterpret_cast<uint8_t*>(full_bar) + 512;   // kEpiWarps x 4 KB epilogue staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);      // leader's copy is the live one: its producer arms it
      mbar_init(&empty_bar[i], 1);     // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * kEpiWarps);   // leader's copy: epilogue warps of BOTH CTAs
    }
    if constexpr (CLC) {
      mbar_init(&feed->clc_bar, 1);
      mbar_init(&feed->peer_armed, 1);
      for (int i = 0; i < 4; ++i) {
        mbar_init(&feed->full[i], 1);
        mbar_init(&feed->empty[i], kEpiWarps + (cluster_ctarank() == 0 ? 1 : 0));   // epilogue warps (+ MMA warp)
      }
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // (p.num_m_blk counts 256-row pair tiles here)

  if (warp < 4) {
  reg_dealloc<56>();          // control warpgroup (TMA, MMA, TMEM allocator, spare): gives its registers away
  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ========================
    int stage = 0;
    uint32_t phase = 0;
    GemmWork work(p, pair, num_pairs);
    int tile, kb0, kb1;
    int unit = pair;          // CLC: this cluster's own unit first
    uint32_t fit = 0;
    while (true) {
      if constexpr (CLC) {
        // publish the unit (or the end marker) to the other roles of this CTA
        const int slot = fit & 3;
        mbar_wait(&feed->empty[slot], ((fit >> 2) & 1) ^ 1);
        if (elect_one_sync()) {
          *reinterpret_cast<volatile int*>(&feed->ring[slot]) = unit;
          mbar_arrive(&feed->full[slot]);
        }
        __syncwarp();
        if (unit < 0) break;
        unit_decode(p, unit, tile, kb0, kb1);
      } else {
        if (!work.next(tile, kb0, kb1)) break;
      }
      const int m_blk = tile / p.num_n_blk, n_blk = tile % p.num_n_blk;
      const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
      const int n0 = n_blk * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
          uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
          uint8_t* b_dst = sB + stage * B_HALF_BYTES;
          if (!p.a_mn) {
            tma_load_2d_2cta(a_dst, &tmA, bar, kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_2cta(a_dst + j * (BK * 128), &tmA, bar, m0 + j * 64, kb * BK);
          }
          if (!p.b_mn) {
            tma_load_2d_2cta(b_dst, &tmB, bar, kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_2cta(b_dst + j * (BK * 128), &tmB, bar, n0 + j * 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      if constexpr (CLC) {
        // next unit: cancel a cluster that has not started yet and take its place.  Every CTA arms its own
        // barrier; the leader queries once the peer has armed (and therefore finished reading the previous
        // response, which the multicast write is about to replace)
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&feed->clc_bar, 16);
          if (rank != 0) mbar_arrive_release_cluster(mapa_shared(smem_u32(&feed->peer_armed), 0));
        }
        __syncwarp();
        if (rank == 0) {
          mbar_wait(&feed->peer_armed, fit & 1);
          if (elect_one_sync()) clc_try_cancel(&feed->clc_resp, &feed->clc_bar);
          __syncwarp();
        }
        mbar_wait(&feed->clc_bar, fit & 1);
        const int first_cta = clc_decode(&feed->clc_resp);
        fence_proxy_async_smem();      // this read is ordered before the next asynchronous write of the response
        unit = first_cta < 0 ? -1 : first_cta >> 1;
        ++fit;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ========================
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16_f32(256, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
      const uint32_t a_kstep = (p.a_mn ? 2048 : 32) >> 4, b_kstep = (p.b_mn ? 2048 : 32) >> 4;
      const uint64_t a_desc0 = umma_smem_desc_sw128(smem_u32(sA), a_lbo, 1024);
      const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(sB), b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      GemmWork work(p, pair, num_pairs);
      int tile, kb0, kb1;
      uint32_t fit = 0;
      while (true) {
        if constexpr (CLC) {
          const int unit = feed_take(feed, fit);
          if (unit < 0) break;
          unit_decode(p, unit, tile, kb0, kb1);
        } else {
          if (!work.next(tile, kb0, kb1)) break;
        }
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t a_desc = a_desc0 + stage * (A_STAGE_BYTES >> 4);
            const uint64_t b_desc = b_desc0 + stage * (B_HALF_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss_2cta(d_tmem, a_desc + k * a_kstep, b_desc + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_2cta(&empty_bar[stage]);                      // frees the slot in both CTAs
            if (kb == kb1 - 1) umma_commit_2cta(&tfull_bar[acc]);     // accumulator complete -> both epilogues
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        if (++acc == 2) acc = 0, acc_phase ^= 1;
      }
    }
  }
  } else {
    reg_alloc<CO ? 208 : 224>();           // the two epilogue warpgroups take them (128 x 56 + 256 x 208 <= 384 x 160)
    // ================================ epilogue (both CTAs, own 128 rows) ==============
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t stage = smem_u32(sStage) + (warp - 4) * kEpiStageBytes;   // shared-space address
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    GemmWork work(p, pair, num_pairs);
    int tile, kb0, kb1;
    uint32_t fit = 0;
    while (true) {
      if constexpr (CLC) {
        const int unit = feed_take(feed, fit);
        if (unit < 0) break;
        unit_decode(p, unit, tile, kb0, kb1);
      } else {
        if (!work.next(tile, kb0, kb1)) break;
      }
      const int m_blk = tile / p.num_n_blk, n_blk = tile % p.num_n_blk;
      const int row0 = m_blk * 256 + static_cast<int>(rank) * 128 + q * 32;
      // in-place RESID (aux_in aliases C): this warp's tile is only ever touched by this warp, and its
      // loads are issued before its stores.
      // ALL auxiliary tiles of this warp (residual / pre-activation / position rows of its BN / 64 chunks) are
      // requested up front, before the accumulator is even complete: with one chunk of look-ahead the warp sat on
      // the global-load latency of every chunk (ncu: the staging store of the aux tile was the top stall, out-proj
      // ran at 2.4 TB/s of epilogue traffic) -- an epilogue warp needs its whole tile's bytes in flight.
      constexpr int NCH = BN / 32 / (kEpiWarps / 4);
      uint4 aux[NCH][8];
      if constexpr (epilogue_has_aux<EPI>()) {
#pragma unroll
        for (int i = 0; i < NCH; ++i)
          epilogue_aux_load<EPI>(p, row0, n_blk * BN + (half + i * (kEpiWarps / 4)) * 32, lane, aux[i]);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = half + i * (kEpiWarps / 4);
        const int col0 = n_blk * BN + c * 32;
        if (col0 < p.N) {  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (row0 < p.M) epilogue_chunk<EPI, OUT_F32>(p, row0, col0, r, aux[i], stage, lane);   // warp-uniform
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader0 + acc * 8);
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
  }

  // neither CTA may leave (or free TMEM) while its partner can still touch its barriers / memory
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, bool OUT_F32, bool CLC, bool CO>
static int launch_gemm2_co(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                           cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  auto kern = gemm_tcgen05_2cta_kernel<BN, EPI, OUT_F32, CLC, CO>;
  static bool configured = false;  // benign race: the attribute call is idempotent
  if (!configured) {
    MISSM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  kern<<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
template <int BN, int EPI, bool OUT_F32, bool CLC>
static int launch_gemm2_sched(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                              cudaStream_t stream) {
  if (coresident()) return launch_gemm2_co<BN, EPI, OUT_F32, CLC, true>(tmA, tmB, p, grid, stream);
  return launch_gemm2_co<BN, EPI, OUT_F32, CLC, false>(tmA, tmB, p, grid, stream);
}
// grid < 0: dynamic schedule, one cluster per work unit (-grid CTAs)
template <int BN, int EPI, bool OUT_F32>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                        cudaStream_t stream) {
  if (grid < 0) return launch_gemm2_sched<BN, EPI, OUT_F32, true>(tmA, tmB, p, -grid, stream);
  return launch_gemm2_sched<BN, EPI, OUT_F32, false>(tmA, tmB, p, grid, stream);
}

template <int BN>
static int dispatch_epi2(int epi, bool out_f32, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p,
                         int grid, cudaStream_t s) {
  switch (epi) {
    case MISSM_EPI_LINEAR:
      return out_f32 ? launch_gemm2<BN, MISSM_EPI_LINEAR, true>(a, b, p, grid, s)
                     : launch_gemm2<BN, MISSM_EPI_LINEAR, false>(a, b, p, grid, s);
    case MISSM_EPI_GELU:
      return launch_gemm2<BN, MISSM_EPI_GELU, false>(a, b, p, grid, s);
    case MISSM_EPI_RESID:
      return launch_gemm2<BN, MISSM_EPI_RESID, true>(a, b, p, grid, s);
    case MISSM_EPI_DGELU:
      return launch_gemm2<BN, MISSM_EPI_DGELU, false>(a, b, p, grid, s);
    case MISSM_EPI_PATCH:
      return launch_gemm2<BN, MISSM_EPI_PATCH, true>(a, b, p, grid, s);
  }
  MISSM_REQUIRE(false, "unknown epilogue %d", epi);
}

// Called by missm_gemm_bf16 (gemm_tcgen05.cu) after argument validation; `p` carries everything but
// the tiling.  Returns -1 if this path declines the shape.
int gemm_launch_2cta(const missm_gemm_args* a, GemmParams p, cudaStream_t stream) {
  const int kPairs = persistent_sms() / 2;
  p.num_m_blk = (a->M + 255) / 256;
  p.num_kblk = (a->K + BK - 1) / BK;
  auto waves_eff = [&](int bn) {
    long tiles = static_cast<long>(p.num_m_blk) * ((a->N + bn - 1) / bn);
    long waves = (tiles + kPairs - 1) / kPairs;
    return static_cast<double>(tiles) * bn / (static_cast<double>(waves) * kPairs * 256.0);
  };
  int bn = 256;
  if (a->force_bn == 128 || a->force_bn == 256)
    bn = a->force_bn;
  else if (a->N <= 128 || waves_eff(128) > 1.15 * waves_eff(256))
    bn = 128;
  p.num_n_blk = (a->N + bn - 1) / bn;

  const long tiles = static_cast<long>(p.num_m_blk) * p.num_n_blk;
  int splits = 1;
  const bool may_split = (a->epilogue == MISSM_EPI_LINEAR && a->out_f32 && a->bias == nullptr &&
                          a->scale_cols == 0 && a->split_k != 1 && a->colsum_out == nullptr && a->colsum_part == nullptr);
  if (may_split) {
    if (a->split_k > 1) {
      splits = a->split_k;
    } else if (a->split_k == 0) {
      // uniform K split s that minimises waves(s) / s, waves(s) = ceil(tiles * s / pairs); only worth the
      // fp32 atomics (measured) when it removes >= 15 % of the time: 48 tiles (qkv wgrad) -> 3, 16 tiles
      // (out-proj wgrad) -> 4, 64 tiles (fc wgrad) -> 1
      double best = 1.0 * ((tiles + kPairs - 1) / kPairs);
      for (int s = 2; s <= 8 && p.num_kblk / s >= 16; ++s) {
        const double c = static_cast<double>((tiles * s + kPairs - 1) / kPairs) / s;
        if (c < 0.85 * best) best = c, splits = s;
      }
    }
    if (splits > p.num_kblk) splits = p.num_kblk;
    if (splits < 1) splits = 1;
  }
  p.kblk_per_split = (p.num_kblk + splits - 1) / splits;
  p.num_splits = (p.num_kblk + p.kblk_per_split - 1) / p.kblk_per_split;
  p.atomic_out = p.num_splits > 1 ? 1 : 0;
  // stream-K (split_k = -1): kept for experiments -- on the wgrad shapes the extra atomic epilogues cost
  // more than the idle SMs they fill (fc wgrad 92 -> 106-115 us)
  p.stream_k = 0, p.sk_per_cta = 0;
  long work = tiles * p.num_splits;
  int grid_pairs = static_cast<int>(work < kPairs ? work : kPairs);
  if (may_split && a->split_k == -1 && p.num_kblk >= 32) {
    const long total = tiles * p.num_kblk;
    p.stream_k = 1, p.atomic_out = 1;
    p.sk_per_cta = (total + kPairs - 1) / kPairs;
    if (p.sk_per_cta < 8) p.sk_per_cta = 8;
    grid_pairs = static_cast<int>((total + p.sk_per_cta - 1) / p.sk_per_cta);
  }
  if (p.atomic_out) {
    MISSM_CHECK_CUDA(cudaMemset2DAsync(a->C, static_cast<size_t>(a->ldc) * 4, 0, static_cast<size_t>(a->N) * 4, a->M,
                                       stream));
  }

  CUtensorMap tmA, tmB;
  int rc;
  if (!p.a_mn)
    rc = make_tmap_2d_bf16(&tmA, a->A, a->K, a->M, a->lda, BK, BM);
  else
    rc = make_tmap_2d_bf16(&tmA, a->A, a->M, a->K, a->lda, 64, BK);
  if (rc) return rc;
  if (!p.b_mn)
    rc = make_tmap_2d_bf16(&tmB, a->B, a->K, a->N, a->ldb, BK, bn / 2);
  else
    rc = make_tmap_2d_bf16(&tmB, a->B, a->N, a->K, a->ldb, 64, BK);
  if (rc) return rc;

  // dynamic (CLC) schedule unless stream-K, a capped grid (the cap exists to keep SMs free) or the A/B switch
  static const bool force_static = getenv("MISSM_GEMM_STATIC") != nullptr;
  const bool dynamic = !force_static && !p.stream_k && kPairs == kNumSMs / 2 && work > grid_pairs;
  const int grid = dynamic ? -2 * static_cast<int>(work) : 2 * grid_pairs;
  if (bn == 256) return dispatch_epi2<256>(a->epilogue, a->out_f32 != 0, tmA, tmB, p, grid, stream);
  return dispatch_epi2<128>(a->epilogue, a->out_f32 != 0, tmA, tmB, p, grid, stream);
}

}  // namespace missm
