// Masked fusion for the default `sum` head, forward and backward, fp32 (HBM/L2-bound, tiny).
//
// Replaces modal_sum.forward (src/model/baseline.py:52-61):
//     for modal: data = Linear_modal(batch[modal]); data[missing_index == code_modal] = 0
//     inputs = sum(data);  out = LayerNorm(inputs)
// One CTA per sample: the embedding row of each modality is staged in shared memory, each warp
// produces Fd/8 outputs with coalesced float4 reads of the weight rows (the weights stay in L2),
// the boolean mask is evaluated once per (sample, modality) and a masked modality costs no reads
// at all; LayerNorm over the Fd sums happens in the same CTA.
//   algorithmic bytes = B*M*P*4 (embeddings) + M*Fd*P*4 (weights, once) + 2*B*Fd*4
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

struct FusionPtrs {
  const float* emb[MISSM_MAX_TOWERS];
  const float* w[MISSM_MAX_TOWERS];
  const float* b[MISSM_MAX_TOWERS];
  float* d_emb[MISSM_MAX_TOWERS];
  float* d_w[MISSM_MAX_TOWERS];
  float* d_b[MISSM_MAX_TOWERS];
  int32_t code[MISSM_MAX_TOWERS];
};

constexpr int kFusionThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < (blockDim.x >> 5)) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

__global__ void __launch_bounds__(kFusionThreads)
fusion_sum_fwd_kernel(FusionPtrs ptrs, int n_mod, const int64_t* __restrict__ missing_index,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ pre, float* __restrict__ out, float* __restrict__ mean_out,
                      float* __restrict__ rstd_out, int P, int Fd, float eps) {
  extern __shared__ float sm[];
  float* s_emb = sm;        // [P]
  float* s_acc = sm + P;    // [Fd]
  __shared__ float red[32];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t mi = missing_index[b];
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) s_acc[f] = 0.f;
  for (int m = 0; m < n_mod; ++m) {
    if (mi == static_cast<int64_t>(ptrs.code[m])) continue;  // block-uniform: masked modality
    __syncthreads();
    const float* e = ptrs.emb[m] + static_cast<long>(b) * P;
    for (int p = threadIdx.x * 4; p < P; p += blockDim.x * 4)
      *reinterpret_cast<float4*>(s_emb + p) = *reinterpret_cast<const float4*>(e + p);
    __syncthreads();
    for (int f = warp; f < Fd; f += nwarps) {
      const float* wr = ptrs.w[m] + static_cast<long>(f) * P;
      float acc = 0.f;
      for (int p = lane * 4; p < P; p += 128) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + p));
        const float4 e4 = *reinterpret_cast<const float4*>(s_emb + p);
        acc += w4.x * e4.x + w4.y * e4.y + w4.z * e4.z + w4.w * e4.w;
      }
      acc = warp_sum(acc);
      if (lane == 0) s_acc[f] += acc + ptrs.b[m][f];
    }
  }
  __syncthreads();
  float part = 0.f;
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) part += s_acc[f];
  const float mean = block_sum(part, red) / Fd;
  part = 0.f;
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) {
    const float d = s_acc[f] - mean;
    part += d * d;
  }
  const float rstd = rsqrtf(block_sum(part, red) / Fd + eps);
  if (threadIdx.x == 0) mean_out[b] = mean, rstd_out[b] = rstd;
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) {
    pre[static_cast<long>(b) * Fd + f] = s_acc[f];
    out[static_cast<long>(b) * Fd + f] = (s_acc[f] - mean) * rstd * gamma[f] + beta[f];
  }
}

// per sample: d_pre = LN'(d_out);  dg_part[b, f] = d_out * xhat,  db_part[b, f] = d_out
__global__ void __launch_bounds__(kFusionThreads)
fusion_sum_bwd_norm_kernel(const float* __restrict__ d_out, const float* __restrict__ pre,
                           const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, float* __restrict__ d_pre,
                           float* __restrict__ dg_part, float* __restrict__ db_part, int Fd) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float mu = mean[b], rs = rstd[b];
  float s1 = 0.f, s2 = 0.f;
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) {
    const long i = static_cast<long>(b) * Fd + f;
    const float xh = (pre[i] - mu) * rs, g = d_out[i] * gamma[f];
    s1 += g, s2 += g * xh;
    dg_part[i] = d_out[i] * xh;
    db_part[i] = d_out[i];
  }
  const float c1 = block_sum(s1, red) / Fd;
  const float c2 = block_sum(s2, red) / Fd;
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) {
    const long i = static_cast<long>(b) * Fd + f;
    const float xh = (pre[i] - mu) * rs, g = d_out[i] * gamma[f];
    d_pre[i] = rs * (g - c1 - xh * c2);
  }
}

// d_emb[m][b, p] = present ? sum_f d_pre[b, f] * W_m[f, p] : 0        grid (B, n_mod)
__global__ void __launch_bounds__(kFusionThreads)
fusion_sum_bwd_emb_kernel(FusionPtrs ptrs, const int64_t* __restrict__ missing_index,
                          const float* __restrict__ d_pre, int P, int Fd) {
  extern __shared__ float s_d[];  // [Fd]
  const int b = blockIdx.x, m = blockIdx.y;
  float* de = ptrs.d_emb[m] + static_cast<long>(b) * P;
  if (missing_index[b] == static_cast<int64_t>(ptrs.code[m])) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) de[p] = 0.f;
    return;
  }
  for (int f = threadIdx.x; f < Fd; f += blockDim.x) s_d[f] = d_pre[static_cast<long>(b) * Fd + f];
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float acc = 0.f;
    const float* w = ptrs.w[m] + p;
#pragma unroll 4
    for (int f = 0; f < Fd; ++f) acc += s_d[f] * __ldg(w + static_cast<long>(f) * P);
    de[p] = acc;
  }
}

// d_w[m][f, p] = sum_b present(b, m) d_pre[b, f] emb_m[b, p];  d_b[m][f] = sum_b present d_pre[b, f]
// grid (Fd, n_mod)
__global__ void __launch_bounds__(kFusionThreads)
fusion_sum_bwd_w_kernel(FusionPtrs ptrs, const int64_t* __restrict__ missing_index,
                        const float* __restrict__ d_pre, int B, int P, int Fd) {
  extern __shared__ float s_g[];  // [B] masked d_pre[:, f]
  const int f = blockIdx.x, m = blockIdx.y;
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    s_g[b] = (missing_index[b] == static_cast<int64_t>(ptrs.code[m])) ? 0.f : d_pre[static_cast<long>(b) * Fd + f];
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float acc = 0.f;
    const float* e = ptrs.emb[m] + p;
    for (int b = 0; b < B; ++b) acc += s_g[b] * e[static_cast<long>(b) * P];
    ptrs.d_w[m][static_cast<long>(f) * P + p] = acc;
  }
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += s_g[b];
    ptrs.d_b[m][f] = acc;
  }
}

__global__ void reduce_rows_f32_kernel(const float* __restrict__ part, int R, int n, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += part[static_cast<long>(r) * n + c];
  out[c] = s;
}

}  // namespace missm

using namespace missm;

static int fill_ptrs(FusionPtrs& p, const missm_fusion_sum_args* a) {
  MISSM_REQUIRE(a->n_modal >= 1 && a->n_modal <= MISSM_MAX_TOWERS, "fusion_sum: n_modal=%d", a->n_modal);
  MISSM_REQUIRE(a->P % 4 == 0 && a->Fd >= 1, "fusion_sum: P=%d Fd=%d", a->P, a->Fd);
  for (int i = 0; i < MISSM_MAX_TOWERS; ++i) {
    const bool on = i < a->n_modal;
    p.emb[i] = on ? a->emb[i] : nullptr;
    p.w[i] = on ? a->weight[i] : nullptr;
    p.b[i] = on ? a->bias[i] : nullptr;
    p.d_emb[i] = on ? a->d_emb[i] : nullptr;
    p.d_w[i] = on ? a->d_weight[i] : nullptr;
    p.d_b[i] = on ? a->d_bias[i] : nullptr;
    p.code[i] = on ? a->codes[i] : -1;
  }
  return 0;
}

extern "C" int missm_fusion_sum_fwd(const missm_fusion_sum_args* a, void* stream) {
  if (a->B == 0) return 0;
  FusionPtrs p;
  if (int rc = fill_ptrs(p, a)) return rc;
  const size_t smem = sizeof(float) * (a->P + a->Fd);
  MISSM_REQUIRE(smem <= 48 * 1024, "fusion_sum: P + Fd too large for shared memory");
  fusion_sum_fwd_kernel<<<a->B, kFusionThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      p, a->n_modal, a->missing_index, a->gamma, a->beta, a->pre, a->out, a->mean, a->rstd, a->P, a->Fd, a->eps); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// workspace: 3 * B * Fd floats (d_pre, dgamma partials, dbeta partials)
extern "C" int missm_fusion_sum_bwd(const missm_fusion_sum_args* a, const float* d_out, float* workspace,
                                    float* d_gamma, float* d_beta, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->B == 0) return 0;
  FusionPtrs p;
  if (int rc = fill_ptrs(p, a)) return rc;
  const long n = static_cast<long>(a->B) * a->Fd;
  float* d_pre = workspace;
  float* dg_part = workspace + n;
  float* db_part = workspace + 2 * n;
  fusion_sum_bwd_norm_kernel<<<a->B, kFusionThreads, 0, st>>>(d_out, a->pre, a->mean, a->rstd, a->gamma, d_pre,
                                                              dg_part, db_part, a->Fd); note_launch();
  reduce_rows_f32_kernel<<<(a->Fd + 127) / 128, 128, 0, st>>>(dg_part, a->B, a->Fd, d_gamma); note_launch();
  reduce_rows_f32_kernel<<<(a->Fd + 127) / 128, 128, 0, st>>>(db_part, a->B, a->Fd, d_beta); note_launch();
  fusion_sum_bwd_emb_kernel<<<dim3(a->B, a->n_modal), kFusionThreads, sizeof(float) * a->Fd, st>>>(
      p, a->missing_index, d_pre, a->P, a->Fd); note_launch();
  MISSM_REQUIRE(sizeof(float) * a->B <= 48 * 1024, "fusion_sum: batch too large for shared memory");
  fusion_sum_bwd_w_kernel<<<dim3(a->Fd, a->n_modal), kFusionThreads, sizeof(float) * a->B, st>>>(
      p, a->missing_index, d_pre, a->B, a->P, a->Fd); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
