// Probe of Blackwell cluster launch control (clusterlaunchcontrol.try_cancel) for clusters of 2:
// grid = 2 * T CTAs; every running cluster keeps cancelling not-yet-launched clusters and takes their
// work.  Checks: every tile id is processed exactly once; reports how many clusters actually ran.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t d;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(d) : "r"(smem_u32(b)), "r"(par) : "memory");
  return d != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
clc_probe(int* visits, int* ran, int spin, long long* lat) {
  __shared__ __align__(16) uint4 resp;
  __shared__ uint64_t bar;
  extern __shared__ uint8_t big[];   // force one CTA per SM
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  int tile = blockIdx.x >> 1;
  if (threadIdx.x == 0 && rank == 0) atomicAdd(ran, 1);
  uint32_t phase = 0;
  int guard = 0;
  while (true) {
    if (threadIdx.x == 0) {
      atomicAdd(&visits[tile * 2 + rank], 1);
      mbar_expect_tx(&bar, 16);                 // every CTA arms its own barrier
    }
    // the pair must not re-query before both CTAs armed / consumed: cluster barrier per tile (probe only)
    cluster_sync();
    if (threadIdx.x == 0 && rank == 0) {
      asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
                   ::"r"(smem_u32(&resp)), "r"(smem_u32(&bar)) : "memory");
    }
    long long t0 = clock64();
    int spins = 0;
    while (!mbar_try(&bar, phase)) { if (++spins > (1 << 22)) { printf("timeout blk %d\n", blockIdx.x); __trap(); } }
    if (threadIdx.x == 0 && rank == 0) { atomicAdd((unsigned long long*)lat, (unsigned long long)(clock64() - t0)); atomicAdd((unsigned long long*)(lat + 1), 1ull); }
    while (clock64() - t0 < spin) {}            // "work"
    phase ^= 1;
    uint32_t valid = 0, x = 0, y, z;
    asm volatile("{\n.reg .pred p1;\n.reg .b128 r;\nld.shared.b128 r, [%4];\n"
                 "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\nselp.u32 %3, 1, 0, p1;\n"
                 "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n}\n"
                 : "=r"(x), "=r"(y), "=r"(z), "=r"(valid) : "r"(smem_u32(&resp)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (!valid) break;
    tile = static_cast<int>(x) >> 1;
    if (++guard > 100000) break;
  }
  cluster_sync();
}

int main() {
  const int T = 1000;
  int *visits, *ran;
  cudaMalloc(&visits, T * 2 * sizeof(int));
  cudaMalloc(&ran, sizeof(int));
  cudaMemset(visits, 0, T * 2 * sizeof(int));
  cudaMemset(ran, 0, sizeof(int));
  cudaFuncSetAttribute(clc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  long long* lat;
  cudaMalloc(&lat, 16);
  clc_probe<<<2 * T, 128, 200 * 1024>>>(visits, ran, 20000, lat);   // warm-up
  cudaDeviceSynchronize();
  cudaMemset(visits, 0, T * 2 * sizeof(int));
  cudaMemset(ran, 0, sizeof(int));
  cudaMemset(lat, 0, 16);
  cudaEventRecord(e0);
  clc_probe<<<2 * T, 128, 200 * 1024>>>(visits, ran, 20000, lat);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<int> h(T * 2);
  int hr = 0;
  cudaMemcpy(h.data(), visits, T * 2 * sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(&hr, ran, sizeof(int), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < T * 2; ++i) bad += (h[i] != 1);
  long long hl[2];
  cudaMemcpy(hl, lat, 16, cudaMemcpyDeviceToHost);
  printf("try_cancel latency (issue -> response visible): %.0f cycles average over %lld queries\n", (double)hl[0] / hl[1], hl[1]);
  printf("err=%s tiles=%d bad=%d clusters_that_ran=%d time=%.3f ms (ideal %.3f ms for %d tiles of 20000 cycles on 74 pairs)\n",
         cudaGetErrorString(err), T, bad, hr, ms, (T + 73) / 74 * 20000 / 1.9e6, T);
  return bad != 0 || err != cudaSuccess;
}
