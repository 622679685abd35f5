// GPU input pipeline for the video and audio modalities (second half).
//
// Video -- replaces the per-sample host chain of languagebind/video/processing_video.py:25-66 applied to the 8 decoded
// frames of a clip (the decord / opencv branches: x / 255 -> NormalizeVideo -> ShortSideScale(224) [pytorchvideo:
// bilinear, align_corners = False, no antialias] -> CenterCropVideo(224) -> RandomHorizontalFlipVideo): one launch
// reads the decoded uint8 frames [T, H, W, 3] once and writes the normalised fp32 clip [3, T, S, S]; only the crop is
// ever computed (normalisation is affine and the resample linear, so they commute).
//
// Audio -- replaces torchaudio.compliance.kaldi.fbank as the reference calls it (audio/processing_audio.py:96-110:
// 25 ms / 10 ms frames, hanning window, 112 mel bins, htk_compat, no dither, no energy) and the chunk / repeat /
// normalise tail of waveform2melspec (:53-94): waveform -> [3, num_mel_bins, target_length] fp32 on the device.
// Per frame: remove DC, pre-emphasis 0.97, window, zero-pad to 512, |DFT|^2 (a direct 257 x 400 DFT per frame from a
// shared twiddle table: 0.2 GFLOP per clip, far below the cost of moving it), mel filterbank, log.
#include "../../include/missm_b200.h"
#include "missm_common.cuh"

namespace missm {

// ------------------------------------------------------------------------------------------------ video
__global__ void __launch_bounds__(256)
video_preprocess_kernel(const missm_video_args a, int RH, int RW, int top, int left, float sy, float sx) {
  const int S = a.S;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (idx >= S * S) return;
  const int oy = idx / S, ox0 = idx % S;
  const int ox = a.hflip ? S - 1 - ox0 : ox0;          // RandomHorizontalFlipVideo: clip.flip(-1)
  // aten upsample_bilinear2d, align_corners = False: src = scale * (dst + 0.5) - 0.5, clamped at 0
  float fy = sy * (static_cast<float>(oy + top) + 0.5f) - 0.5f;
  float fx = sx * (static_cast<float>(ox + left) + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy, fx = fx < 0.f ? 0.f : fx;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + (y0 < a.H - 1 ? 1 : 0), x1 = x0 + (x0 < a.W - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const uint8_t* f = static_cast<const uint8_t*>(a.src) + static_cast<long>(t) * a.H * a.W * 3;
  const uint8_t* p00 = f + (static_cast<long>(y0) * a.W + x0) * 3;
  const uint8_t* p01 = f + (static_cast<long>(y0) * a.W + x1) * 3;
  const uint8_t* p10 = f + (static_cast<long>(y1) * a.W + x0) * 3;
  const uint8_t* p11 = f + (static_cast<long>(y1) * a.W + x1) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // (x / 255 - mean) / std per source pixel, then the 4-tap blend, in the reference's order of operations
    const float v00 = (p00[c] / 255.0f - a.mean[c]) / a.std_[c], v01 = (p01[c] / 255.0f - a.mean[c]) / a.std_[c];
    const float v10 = (p10[c] / 255.0f - a.mean[c]) / a.std_[c], v11 = (p11[c] / 255.0f - a.mean[c]) / a.std_[c];
    const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
    a.dst[((static_cast<long>(c) * a.T + t) * S + oy) * S + ox0] = v;
  }
}

// ------------------------------------------------------------------------------------------------ audio
constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_PAD = 512, FB_BINS = FB_PAD / 2 + 1;

__global__ void wave_sum_kernel(const float* __restrict__ w, long n, double* __restrict__ out) {
  double s = 0.0;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    s += w[i];
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

// one CTA (256 threads) per frame
__global__ void __launch_bounds__(256)
fbank_kernel(const float* __restrict__ wave, const double* __restrict__ wave_sum, long n_total, int n_frames,
             const float* __restrict__ mel_w /* [n_mel, FB_BINS] */, int n_mel, float* __restrict__ mel_out /* [n_frames, n_mel] */) {
  __shared__ float fr[FB_PAD];
  __shared__ float tc[FB_PAD], ts[FB_PAD];
  __shared__ float pw[FB_BINS + 3];
  __shared__ float red[8];
  const int f = blockIdx.x, tid = threadIdx.x;
  if (f >= n_frames) return;
  const float mean_all = static_cast<float>(wave_sum[0] / static_cast<double>(n_total));    // audio_data -= audio_data.mean()
  const float* x = wave + static_cast<long>(f) * FB_SHIFT;
  for (int i = tid; i < FB_PAD; i += 256) {
    float s, c;
    sincospif(2.0f * i / FB_PAD, &s, &c);
    tc[i] = c, ts[i] = s;
    fr[i] = i < FB_WIN ? x[i] - mean_all : 0.f;
  }
  __syncthreads();
  // remove_dc_offset: subtract the frame's own mean
  float part = 0.f;
  for (int i = tid; i < FB_WIN; i += 256) part += fr[i];
  part = warp_sum(part);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  float fm = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) fm += red[i];
  fm *= (1.0f / FB_WIN);
  // pre-emphasis (previous sample, the first one replicated) and the hanning window, out of place
  float v0 = 0.f, v1 = 0.f;
  {
    const int i0 = tid, i1 = tid + 256;
    const float cur0 = fr[i0] - fm, prev0 = fr[i0 > 0 ? i0 - 1 : 0] - fm;
    v0 = (cur0 - 0.97f * prev0) * (0.5f - 0.5f * cospif(2.0f * i0 / (FB_WIN - 1)));
    if (i1 < FB_WIN) {
      const float cur1 = fr[i1] - fm, prev1 = fr[i1 - 1] - fm;
      v1 = (cur1 - 0.97f * prev1) * (0.5f - 0.5f * cospif(2.0f * i1 / (FB_WIN - 1)));
    }
  }
  __syncthreads();
  fr[tid] = v0;
  if (tid + 256 < FB_WIN) fr[tid + 256] = v1;
  __syncthreads();
  // power spectrum: bins tid (and 256 for thread 0)
  for (int k = tid; k < FB_BINS; k += 256) {
    float re = 0.f, im = 0.f;
#pragma unroll 4
    for (int n = 0; n < FB_WIN; ++n) {
      const int j = (k * n) & (FB_PAD - 1);
      re = fmaf(fr[n], tc[j], re), im = fmaf(fr[n], ts[j], im);
    }
    pw[k] = re * re + im * im;
  }
  __syncthreads();
  for (int m = tid; m < n_mel; m += 256) {
    const float* wrow = mel_w + static_cast<long>(m) * FB_BINS;
    float e = 0.f;
    for (int k = 0; k < FB_BINS; ++k) e = fmaf(__ldg(wrow + k), pw[k], e);
    mel_out[static_cast<long>(f) * n_mel + m] = __logf(fmaxf(e, 1.1920928955078125e-07f));
  }
}

// out[c, m, t] = (mel[(off[c] + t) % n_frames, m] - mean) / (2 std)      [3, n_mel, target]
__global__ void mel_pack_kernel(const float* __restrict__ mel, int n_frames, int n_mel, int target, int o0, int o1, int o2,
                                float mean, float inv_2std, float* __restrict__ out) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long total = 3L * n_mel * target;
  if (idx >= total) return;
  const int t = static_cast<int>(idx % target);
  const int m = static_cast<int>((idx / target) % n_mel);
  const int c = static_cast<int>(idx / (static_cast<long>(target) * n_mel));
  const int off = c == 0 ? o0 : (c == 1 ? o1 : o2);
  out[idx] = (mel[static_cast<long>((off + t) % n_frames) * n_mel + m] - mean) * inv_2std;
}

}  // namespace missm

using namespace missm;

extern "C" int missm_video_preprocess(const missm_video_args* a, void* stream) {
  MISSM_REQUIRE(a && a->src && a->dst && a->T > 0 && a->H > 0 && a->W > 0 && a->S > 0, "video_preprocess: bad arguments");
  // pytorchvideo.transforms.functional.short_side_scale + torchvision center_crop
  int RH, RW;
  if (a->W < a->H) RH = static_cast<int>(floor((static_cast<double>(a->H) / a->W) * a->S)), RW = a->S;
  else RH = a->S, RW = static_cast<int>(floor((static_cast<double>(a->W) / a->H) * a->S));
  MISSM_REQUIRE(RH >= a->S && RW >= a->S, "video_preprocess: %d x %d after scaling is smaller than the crop", RH, RW);
  const int top = static_cast<int>(nearbyint((RH - a->S) / 2.0)), left = static_cast<int>(nearbyint((RW - a->S) / 2.0));
  const float sy = static_cast<float>(a->H) / RH, sx = static_cast<float>(a->W) / RW;       // aten: input / output size
  dim3 grid((a->S * a->S + 255) / 256, a->T);
  video_preprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a, RH, RW, top, left, sy, sx); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int missm_fbank_num_frames(int64_t n_samples) {
  return n_samples < FB_WIN ? 0 : static_cast<int>(1 + (n_samples - FB_WIN) / FB_SHIFT);
}

extern "C" int missm_audio_fbank(const missm_fbank_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MISSM_REQUIRE(a && a->wave && a->mel_weights && a->mel && a->out && a->wave_sum, "audio_fbank: null pointer");
  MISSM_REQUIRE(a->n_mel > 0 && a->target > 0 && a->n_total >= a->n_samples, "audio_fbank: bad sizes");
  const int nf = missm_fbank_num_frames(a->n_samples);
  MISSM_REQUIRE(nf > 0, "audio_fbank: %ld samples are shorter than one 25 ms frame", (long)a->n_samples);
  MISSM_CHECK_CUDA(cudaMemsetAsync(a->wave_sum, 0, sizeof(double), st));
  wave_sum_kernel<<<64, 256, 0, st>>>(a->wave_all ? a->wave_all : a->wave, a->n_total, a->wave_sum); note_launch();
  fbank_kernel<<<nf, 256, 0, st>>>(a->wave, a->wave_sum, a->n_total, nf, a->mel_weights, a->n_mel, a->mel); note_launch();
  const long total = 3L * a->n_mel * a->target;
  mel_pack_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(a->mel, nf, a->n_mel, a->target, a->offsets[0], a->offsets[1],
                                                                          a->offsets[2], a->mean, 1.0f / (2.0f * a->std_), a->out); note_launch();
  MISSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
