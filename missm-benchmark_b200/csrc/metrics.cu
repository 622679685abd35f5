// Evaluation metrics accumulated ON THE DEVICE.  The reference's evaluate()
// (train_ddp.py:88-133, test.py:21-66) takes, per batch, argmax + softmax of the logits and the batch's
// CrossEntropy, and pulls all three to the host every batch (`loss.item()`, `.cpu().numpy()`: two host
// synchronisations per batch in a loop whose forward pass is otherwise asynchronous); accuracy / macro-F1 come from
// sklearn at the end.  Here one launch per batch folds the batch into device-resident accumulators and nothing
// synchronises until the epoch's read-out:
//   confusion[label, pred] += 1      (int64 [C, C]; pred = FIRST maximum, as torch.argmax)
//   loss_sum += mean over the batch of  logsumexp(logits) - logits[label]      (what criterion(outputs, labels).item()
//                                        adds per batch for the default CrossEntropyLoss(reduction='mean'))
//   probs[b, :] = softmax(logits[b, :])  (kept for roc_auc_score, which needs every sample's scores)
// One warp per sample, C <= 1024 classes.  HBM / latency-bound: B * C * 4 bytes in, the same out.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

__global__ void eval_accumulate_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int B, int C,
                                       float* __restrict__ probs, long long* __restrict__ confusion,
                                       double* __restrict__ loss_sum, long long* __restrict__ n_seen) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* row = logits + static_cast<long>(b) * C;
  float mx = -INFINITY;
  int arg = C;
  for (int c = lane; c < C; c += 32) {
    const float v = row[c];
    if (v > mx) mx = v, arg = c;          // strict '>' keeps the first maximum of this lane's columns
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) mx = om, arg = oa;
  }
  float z = 0.f;
  for (int c = lane; c < C; c += 32) z += __expf(row[c] - mx);
  z = warp_sum(z);
  for (int c = lane; c < C; c += 32) probs[static_cast<long>(b) * C + c] = __expf(row[c] - mx) / z;
  if (lane == 0) {
    const int y = static_cast<int>(labels[b]);
    if (y >= 0 && y < C) {
      atomicAdd(reinterpret_cast<unsigned long long*>(confusion + static_cast<long>(y) * C + arg), 1ull);
      atomicAdd(loss_sum, static_cast<double>((mx + __logf(z)) - row[y]) / B);
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(n_seen), 1ull);
  }
}

}  // namespace missm

extern "C" int missm_eval_accumulate(const float* logits, const int64_t* labels, int32_t B, int32_t C, float* probs,
                                     int64_t* confusion, double* loss_sum, int64_t* n_seen, void* stream) {
  using namespace missm;
  if (B == 0) return 0;
  MISSM_REQUIRE(C >= 1 && C <= 1024, "eval_accumulate: C=%d", C);
  MISSM_REQUIRE(logits && labels && probs && confusion && loss_sum && n_seen, "eval_accumulate: null pointer");
  const int warps = 8;
  eval_accumulate_kernel<<<(B + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, labels, B, C, probs, reinterpret_cast<long long*>(confusion), loss_sum, reinterpret_cast<long long*>(n_seen)); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
