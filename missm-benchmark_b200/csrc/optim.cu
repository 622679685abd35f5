// Multi-tensor Adam for the parameters of the towers and the fusion head.
//
// Replaces: `optimizer.step()` of torch.optim.Adam as the reference drives it (train_ddp.py:205 constructs
// Adam(model.parameters(), lr, weight_decay), :254 steps it) -- L2 weight decay folded into the gradient, no amsgrad.
// One launch updates EVERY tensor of a parameter group: the tensors are cut into fixed-size chunks listed in a
// device table, a persistent grid walks the chunk list.  HBM-bound: 16 B read (p, g, m, v) + 12 B written (p, m, v)
// per element = 28 B, + 2 B when the bf16 GEMM-operand copy of the updated weight is emitted in the same pass
// (saves the separate 6 B / element cast kernel of the next forward), + 4 B with the fused zero_grad.
//
// Arithmetic follows torch's single-tensor Adam operation by operation (lerp, mul + addcmul, sqrt / bc2_sqrt + eps,
// addcdiv) in IEEE fp32, so a run can switch optimizers mid-training; results agree to fp32 rounding.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

struct AdamHyper {
  float beta2, om_beta1, om_beta2, eps, weight_decay;   // 1 - beta computed in double on the host, as torch does
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamHyper& h, float step_size,
                                          float bc2_sqrt) {
  if (h.weight_decay != 0.0f) g = fmaf(h.weight_decay, p, g);       // grad.add(param, alpha = weight_decay)
  m = fmaf(h.om_beta1, g - m, m);                                    // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(h.om_beta2 * g, g, v * h.beta2);                        // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + h.eps;                  // (sqrt(v) / bias_correction2_sqrt).add_(eps)
  p = fmaf(-step_size, m / denom, p);                               // param.addcdiv_(exp_avg, denom, -step_size)
}

constexpr int kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(float* const* __restrict__ params, float* const* __restrict__ grads, float* const* __restrict__ exp_avg,
                  float* const* __restrict__ exp_avg_sq, __nv_bfloat16* const* __restrict__ bf16_out,
                  const int64_t* __restrict__ numel, const float* __restrict__ step_size,
                  const float* __restrict__ bc2_sqrt, const int32_t* __restrict__ chunk_tensor,
                  const int64_t* __restrict__ chunk_offset, int n_chunks, long chunk_elems, AdamHyper h, int zero_grads) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int t = chunk_tensor[c];
    float* g = grads[t];
    if (g == nullptr) continue;                                     // no gradient this step: torch skips the tensor
    const long off = chunk_offset[c];
    long n = numel[t] - off;
    if (n > chunk_elems) n = chunk_elems;
    float* p = params[t] + off;
    float* m = exp_avg[t] + off;
    float* v = exp_avg_sq[t] + off;
    g += off;
    __nv_bfloat16* w16 = (bf16_out != nullptr && bf16_out[t] != nullptr) ? bf16_out[t] + off : nullptr;
    const float ss = step_size[t], b2 = bc2_sqrt[t];
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                     (w16 == nullptr || (reinterpret_cast<uintptr_t>(w16) & 7) == 0);
    long done = 0;
    if (vec) {
      const long n4 = n >> 2;
      float4* p4 = reinterpret_cast<float4*>(p);
      float4* g4 = reinterpret_cast<float4*>(g);
      float4* m4 = reinterpret_cast<float4*>(m);
      float4* v4 = reinterpret_cast<float4*>(v);
      // two independent float4 quadruples per thread and iteration: 8 x 16 B loads in flight
      for (long i = threadIdx.x; i < n4; i += 2 * kAdamThreads) {
        const long j = i + kAdamThreads;
        const bool two = j < n4;
        float4 pa = p4[i], ga = __ldcs(g4 + i), ma = m4[i], va = v4[i];
        float4 pb, gb, mb, vb;
        if (two) pb = p4[j], gb = __ldcs(g4 + j), mb = m4[j], vb = v4[j];
        adam_elem(pa.x, ga.x, ma.x, va.x, h, ss, b2);
        adam_elem(pa.y, ga.y, ma.y, va.y, h, ss, b2);
        adam_elem(pa.z, ga.z, ma.z, va.z, h, ss, b2);
        adam_elem(pa.w, ga.w, ma.w, va.w, h, ss, b2);
        p4[i] = pa, m4[i] = ma, v4[i] = va;
        if (zero_grads) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w16 != nullptr) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(pa.x, pa.y), hi = __floats2bfloat162_rn(pa.z, pa.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo), pk.y = *reinterpret_cast<uint32_t*>(&hi);
          reinterpret_cast<uint2*>(w16)[i] = pk;
        }
        if (two) {
          adam_elem(pb.x, gb.x, mb.x, vb.x, h, ss, b2);
          adam_elem(pb.y, gb.y, mb.y, vb.y, h, ss, b2);
          adam_elem(pb.z, gb.z, mb.z, vb.z, h, ss, b2);
          adam_elem(pb.w, gb.w, mb.w, vb.w, h, ss, b2);
          p4[j] = pb, m4[j] = mb, v4[j] = vb;
          if (zero_grads) g4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (w16 != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pb.x, pb.y), hi = __floats2bfloat162_rn(pb.z, pb.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo), pk.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(w16)[j] = pk;
          }
        }
      }
      done = n4 << 2;
    }
    for (long i = done + threadIdx.x; i < n; i += kAdamThreads) {     // unaligned tensors and the < 4 element tail
      float pe = p[i], me = m[i], ve = v[i];
      adam_elem(pe, g[i], me, ve, h, ss, b2);
      p[i] = pe, m[i] = me, v[i] = ve;
      if (zero_grads) g[i] = 0.f;
      if (w16 != nullptr) w16[i] = __float2bfloat16_rn(pe);
    }
  }
}

}  // namespace missm

extern "C" int missm_adam_multi(const missm_adam_args* a, void* stream) {
  using namespace missm;
  MISSM_REQUIRE(a != nullptr, "adam: null args");
  if (a->n_chunks == 0 || a->n_tensors == 0) return 0;
  MISSM_REQUIRE(a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->numel && a->step_size && a->bc2_sqrt &&
                    a->chunk_tensor && a->chunk_offset,
                "adam: null table pointer");
  MISSM_REQUIRE(a->chunk_elems > 0 && a->chunk_elems % 4 == 0, "adam: chunk_elems=%ld must be a positive multiple of 4",
                (long)a->chunk_elems);
  MISSM_REQUIRE(a->beta1 >= 0. && a->beta1 < 1. && a->beta2 >= 0. && a->beta2 < 1. && a->eps >= 0.f,
                "adam: bad hyper-parameters beta1=%f beta2=%f eps=%g", a->beta1, a->beta2, a->eps);
  AdamHyper h{static_cast<float>(a->beta2), static_cast<float>(1.0 - a->beta1), static_cast<float>(1.0 - a->beta2),
              a->eps, a->weight_decay};
  // HBM-bound streaming kernel: 8 CTAs of 256 threads per SM keep ~128 KB of loads in flight per SM
  const int cap = 8 * kNumSMs;
  const int grid = a->n_chunks < cap ? a->n_chunks : cap;
  adam_multi_kernel<<<grid, kAdamThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float* const*>(a->params), reinterpret_cast<float* const*>(a->grads),
      reinterpret_cast<float* const*>(a->exp_avg), reinterpret_cast<float* const*>(a->exp_avg_sq),
      reinterpret_cast<__nv_bfloat16* const*>(a->bf16_out), a->numel, a->step_size, a->bc2_sqrt, a->chunk_tensor,
      a->chunk_offset, a->n_chunks, static_cast<long>(a->chunk_elems), h, a->zero_grads); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
