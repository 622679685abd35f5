// Common device-side PTX wrappers for sm_100a (mbarrier, TMA, tcgen05/TMEM) and host helpers.
// Everything here is written against the PTX ISA directly; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace missm {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
// SMs the persistent kernels (GEMM, tcgen05 attention) spread over.  Default: all 148.  Under data
// parallelism MISSM_PERSISTENT_SMS (even, e.g. 132) leaves a few SMs to NCCL's all-reduce CTAs, which
// cannot co-reside with one-CTA-per-SM kernels that own the whole register file and shared memory.
int persistent_sms();
// one call per kernel launch this library issues (missm_launch_count: bench.py's `gpu_launches`)
void note_launch();

// ----------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define MISSM_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::missm::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                              cudaGetErrorString(_e));                                      \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)
#define MISSM_REQUIRE(cond, ...)                                                            \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ::missm::set_last_error(__VA_ARGS__);                                                 \
      return 2;                                                                             \
    }                                                                                       \
  } while (0)

// ----------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// Explicit shared-space 128-bit accesses.  Pointers carved out of the dynamic shared buffer through
// integer arithmetic lose their address space, and ptxas then emits GENERIC LD/ST (plus a heavier
// fence at every __syncwarp) -- measured 2-3x slower epilogues; these keep LDS/STS.
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// One lane of a CONVERGED warp.  Single-thread work (tcgen05.mma / TMA issue) is written as
// `if (elect_one_sync()) {...}` inside warp-uniform code rather than under `if (lane == 0)`: only
// then does ptxas keep descriptors and TMEM addresses in uniform registers instead of wrapping every
// UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST + branch "waterfall" (~100 issue cycles each).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// QuickGELU (transformers ACT2FN["quick_gelu"]): x * sigmoid(1.702 x).
// sigmoid(y) = 0.5 + 0.5 tanh(y/2): ONE MUFU op (tanh.approx.f32, rel. error 2^-11, far below the
// bf16 rounding of the stored result) instead of ex2 + rcp -- the GEMM epilogues that apply it
// run 32 K elements per tile and were MUFU/issue bound.
__device__ __forceinline__ float fast_sigmoid_1702(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float quick_gelu(float x) { return x * fast_sigmoid_1702(x); }
__device__ __forceinline__ float quick_gelu_grad(float x) {
  const float s = fast_sigmoid_1702(x);
  return fmaf(1.702f * x * s, 1.0f - s, s);
}

// sum over the 32 lanes of v[c] for every column c: butterfly "transpose-reduce" (31 shuffles);
// returns in lane l the total of column l.  Destroys v.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Warp-collective coalesced store of a 32 x 32 bf16 tile whose rows live one per thread (the layout
// tcgen05.ld hands out): rows go through a private 2 KB shared staging tile (64-byte pitch, 16-byte
// chunks XOR-swizzled -> conflict-free both ways) and leave as 16-byte stores covering 8 rows x 64
// contiguous bytes per instruction, instead of 32 different 128-byte lines per instruction.
//   v: this thread's row (lane = row), g: global address of element (row 0, col 0), ld in elements,
//   rows >= nrows are not written.
__device__ __forceinline__ uint32_t stage64_off(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }
__device__ __forceinline__ void warp_store_tile32_bf16(uint32_t stage, int lane, const float (&v)[32],
                                                       __nv_bfloat16* g, long ld, int nrows) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 q;
    q.x = pack_bf16x2(v[8 * c], v[8 * c + 1]), q.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
    q.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), q.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
    sts128(stage + stage64_off(lane, c), q);
  }
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    if (r < nrows) *reinterpret_cast<uint4*>(g + r * ld + c * 8) = lds128(stage + stage64_off(r, c));
  }
  __syncwarp();
}

// Register rebalancing between warpgroups (setmaxnreg): a kernel launched with 384 threads gets 65536 / 384 = 170
// registers per thread; the producer / MMA warps need ~40, so the epilogue warpgroups can take up to 232 and keep
// every auxiliary tile of a warp in flight at once.  Must be executed by ALL four warps of a warpgroup.
//
// Co-residency budget: the persistent one-CTA-per-SM kernels are compiled for kCoResidentRegs registers per thread
// (__maxnreg__), i.e. 384 x 160 = 61 440 of the SM's 65 536 registers, and rebalance INSIDE that pool.  The 4 096
// registers left (and 128 of the 2 048 thread slots, > 1 KB of shared memory) admit one small foreign CTA per SM at any
// time -- the per-parameter gradient copies of torch DDP's reducer and other 128-thread elementwise kernels -- which
// otherwise could only start in the gaps between two persistent kernels: measured at 2 x B200, the reducer's stream
// advanced ~12 launches per ms against ~18 needed, a third of the all-reduce buckets queued up behind it and ran
// after the backward had finished (profiles/r03c_*).
constexpr int kCoResidentRegs = 160;
// Every persistent kernel exists in both variants (template parameter CO): the co-resident one costs ~1.6 % of a
// single-GPU step (fewer registers for the epilogue / softmax warps), so it is selected at run time -- by
// missm_set_coresident(1), which the encoder bank calls when torch.distributed runs more than one rank.
constexpr int kFullRegs = 168;          // 65 536 / 384 threads, rounded down to the allocation unit
bool coresident();                      // gemm_tcgen05.cu
#define MISSM_PERSISTENT_BOUNDS(CO) __maxnreg__((CO) ? kCoResidentRegs : kFullRegs)
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
#ifndef MISSM_MBAR_SPINS
#define MISSM_MBAR_SPINS (1u << 23)   // try_wait itself suspends for a while; this is seconds, kernels take < 1 ms
#endif
#ifndef MISSM_KERNEL_TAG
#define MISSM_KERNEL_TAG "?"
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > MISSM_MBAR_SPINS) {
      printf("missm: mbarrier wait timeout in %s (block %d of %d, thread %d, barrier @smem+%u, parity %u)\n",
             MISSM_KERNEL_TAG, blockIdx.x, gridDim.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs with fp32 accumulate
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05 async ops of this thread complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 2^x on the MUFU pipe (ex2.approx.ftz: -inf -> +0, no denormal fix-up code around it)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ----------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one MMA spans both SMs' tensor cores, each CTA
// feeds its own 128 rows of A and HALF of the B tile from its own shared memory
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope), as CUTLASS' ClusterBarrier::arrive(cta_id): an explicit
  // .release.cluster makes ptxas emit MEMBAR.ALL.GPU, which stalls the epilogue warp until all of its
  // global stores have drained -- the barrier only has to order this warp's TMEM reads (tcgen05 fence)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread retire) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// TMA load into THIS CTA's shared memory, completing transaction bytes on the mbarrier at
// `bar_cluster_addr` (a shared::cluster address: the pair leader's barrier)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// UMMA descriptors (PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor")
// ----------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128B swizzle, version 1 (Blackwell).
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4     bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__host__ __device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr,
                                                                  uint32_t lbo_bytes,
                                                                  uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
//   [4,6) c_format (1 = f32)  [7,10) a_format (1 = bf16)  [10,13) b_format (1 = bf16)
//   [15] a_major (0 = K, 1 = MN)  [16] b_major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// host: TMA tensor-map encoding through the driver entry point (no link-time libcuda dep)
// ----------------------------------------------------------------------------------------
// 2-D bf16 tensor, `inner` contiguous elements, `outer` rows with `ld` elements between rows;
// box = box_inner x box_outer, 128B swizzle (box_inner * 2 bytes must be <= 128).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer);
// 3-D bf16 tensor (d0 contiguous; d1, d2 with element strides), box = b0 x b1 x 1, 128B swizzle
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t b0, uint32_t b1);

}  // namespace missm
