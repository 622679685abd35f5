// How long does a failed mbarrier.try_wait take?  (calibrates the bounded-wait trap of missm_common.cuh)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void probe(long long* out, int spins) {
  __shared__ uint64_t bar;
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
  __syncthreads();
  long long t0 = clock64();
  int n = 0;
  for (int i = 0; i < spins; ++i) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(a), "r"(0u) : "memory");
    n += done;
  }
  if (threadIdx.x == 0) out[0] = clock64() - t0, out[1] = n;
}
int main() {
  long long* d; long long h[2];
  cudaMalloc(&d, 16);
  for (int spins : {100, 1000, 10000}) {
    probe<<<1, 32>>>(d, spins);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("spins %d: %lld cycles total, %.1f cycles per failed try_wait (err %s)\n", spins, h[0], double(h[0]) / spins, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
