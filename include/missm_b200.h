/* missm_b200 -- C ABI of the B200-native (sm_100a) hot path of MissM-Benchmark.
 *
 * The reference (Fieldhunter/MissM-Benchmark) is pure Python and has no FFI of its own; this
 * header is the boundary a maintainer binds with ctypes.  Each entry
 * point cites the reference code whose arithmetic it replaces (file:line under the reference
 * tree; "transformers 4.3x" = the unvendored third-party CLIP modules it imports at
 * languagebind/image/modeling_image.py:11-12).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, nothing allocates
 *     device memory, there is no global mutable state besides the last-error string
 *     (thread-local) and the persistent-grid size (missm_set_persistent_sms), so calls are
 *     re-entrant (autograd worker threads);
 *   - return value 0 = ok; non-zero = error, text in missm_last_error();
 *   - bf16 tensors are row-major with the stated leading dimension in ELEMENTS.
 */
#ifndef MISSM_B200_H_
#define MISSM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MISSM_ABI_VERSION 9   /* 3: missm_set_persistent_sms; 4: fp32 verification mode; 5: missm_adam_multi;
                                 6: missm_image_preprocess; 7: residual-block drivers, launch counter, GEMM profile;
                                 8: missm_patch_embed_implicit; 9: missm_gemm_args.colsum_part */

int missm_version(void);
const char* missm_last_error(void);

/* Select the co-resident variants of the persistent kernels (tcgen05 GEMM, attention forward / backward): compiled for
 * 160 instead of 168 registers per thread, so that one 128-thread, <= 32-register foreign CTA (torch DDP's per-parameter
 * gradient copies) fits on every SM while they run.  Costs ~1.6 % of a single-GPU step, gains more than that under
 * DDP; the encoder bank turns it on when torch.distributed runs more than one rank.  MISSM_CORESIDENT=0/1 in the
 * environment pins it. */
int missm_set_coresident(int32_t on);
/* SMs the persistent kernels (tcgen05 GEMM, tcgen05 attention: one CTA per SM) spread over from now on; n <= 0
 * restores the default (148, or MISSM_PERSISTENT_SMS).  Process-wide, read at every launch.  The host side lowers it
 * for the backward pass under data parallelism so that NCCL's all-reduce CTAs (issued by the unchanged script's DDP
 * wrapper, train_ddp.py:189) find free SMs instead of queueing behind kernels that own whole SMs. */
int missm_set_persistent_sms(int32_t n);

/* ---------------------------------------------------------------------------------------
 * tcgen05 GEMM:  C[m,n] = epilogue( sum_k A[m,k] * B[n,k] )     bf16 in, fp32 accumulate
 *
 * Replaces every nn.Linear on the path (transformers 4.3x CLIPAttention q/k/v/out_proj and
 * CLIPMLP fc1/fc2, called at languagebind/image/modeling_image.py:121,133,140,150; the
 * projection at languagebind/__init__.py:79) and their autograd dgrad/wgrad twins, and the
 * patch-embedding Conv2d (video/modeling_video.py:29-35,45) as a GEMM over patch rows.
 *
 *   a_mn = 0: A stored [M,K] (K contiguous, lda);  a_mn = 1: A stored [K,M] (M contiguous)
 *   b_mn = 0: B stored [N,K] (K contiguous, ldb);  b_mn = 1: B stored [K,N] (N contiguous)
 * so y = x W^T is (a_mn=0,b_mn=0), dgrad dX = dY W is (0,1) on the same W, and
 * wgrad dW = dY^T X is (1,1) -- no operand is ever transposed in memory.
 * ------------------------------------------------------------------------------------- */
enum missm_epilogue {
  MISSM_EPI_LINEAR = 0, /* C = (acc + bias[n]) * (n < scale_cols ? col_scale : 1)          */
  MISSM_EPI_GELU = 1,   /* u = acc + bias; aux_out = bf16(u); C = bf16(u * sigmoid(1.702u)) */
  MISSM_EPI_RESID = 2,  /* C(f32) = aux_in(f32)[m,n] + acc + bias[n]   (may alias C)        */
  MISSM_EPI_DGELU = 3,  /* C = bf16(acc * quickgelu'(aux_in(bf16)[m,n]))                    */
  MISSM_EPI_PATCH = 4   /* C(f32)[(m/P)*(P+1)+1+m%P, n] = acc + aux_in(f32)[1+m%P, n]       */
};

typedef struct missm_gemm_args {
  const void* A; /* bf16 */
  const void* B; /* bf16 */
  void* C;       /* bf16 or f32 (out_f32) */
  const float* bias;  /* [N] or NULL */
  const void* aux_in; /* see epilogue */
  void* aux_out;      /* see epilogue */
  int32_t M, N, K;
  int32_t lda, ldb, ldc;
  int32_t ld_aux_in, ld_aux_out;
  int32_t a_mn, b_mn;
  int32_t epilogue;
  int32_t out_f32;    /* 1: C is float32, 0: C is bf16 */
  int32_t scale_cols; /* LINEAR: columns [0,scale_cols) are multiplied by col_scale */
  float col_scale;
  int32_t patch_P;    /* PATCH: patches per sample */
  int32_t split_k;    /* 0 = auto (only LINEAR + out_f32 + no bias may split), 1 = never, n > 1 = n ways, -1 = stream-K */
  int32_t force_bn;   /* 0 = auto, 128 or 256 = force tile N */
  float* colsum_out;  /* optional [N], PRE-ZEROED: += column sums over m of the fp32 value that is written
                         to C (bias gradient of the consumer layer, fused into the producing GEMM) */
  float* colsum_part; /* optional [missm_gemm_colsum_rows(M), N] (v9): row r receives the column sums of C rows
                         [32 r, 32 r + 32) -- of the values as STORED (rounded to bf16 when C is bf16) -- written with
                         plain stores by the epilogue warp that owns those rows: deterministic, no atomics, no zeroing.
                         Reduce the rows with missm_reduce_partials.  Not with split-K. */
} missm_gemm_args;

int missm_gemm_bf16(const missm_gemm_args* args, void* stream);
int missm_gemm_colsum_rows(int32_t M);   /* rows of a colsum_part workspace: ceil(M / 32) */

/* ---------------------------------------------------------------------------------------
 * Fused attention (head_dim 64).  Replaces transformers 4.3x CLIPAttention's score/softmax/
 * value chain, called at languagebind/image/modeling_image.py:121 (temporal) and :140 (spatial;
 * text adds the causal mask of :441-455 and the padding mask of :500-502).
 * Sequence s, token t lives at row
 *     (s / s_in) * seq_outer + (s % s_in) * seq_inner + t * tok_stride
 * of qkv [rows, ld_qkv] (q | k | v column blocks of width D, q pre-scaled) and of out/d_out
 * [rows, ld_o].  key_mask (int64, 1 = attend, [*, N]) row = mask_rows ? mask_rows[s / mask_div]
 * : s / mask_div.  lse / delta: float [n_seq, H, N].
 * ------------------------------------------------------------------------------------- */
typedef struct missm_attn_args {
  const void* qkv;
  void* out;
  float* lse;
  const void* d_out; /* bwd */
  float* delta;      /* bwd workspace */
  void* dqkv;        /* bwd out, same layout as qkv */
  const int64_t* key_mask;
  const int32_t* mask_rows;
  int64_t ld_qkv, ld_o;
  int64_t seq_outer, seq_inner, tok_stride;
  int32_t D, H, N, head_dim;
  int32_t n_seq, s_in;
  int32_t causal, mask_div;
  float q_scale; /* bwd: dq is multiplied by this */
  float* dqkv_colsum;     /* bwd, optional [3D], PRE-ZEROED: += column sums of dqkv (the q/k/v bias gradients) */
  int32_t colsum_done;    /* bwd, OUT: 1 if the kernels filled dqkv_colsum, 0 if the caller has to reduce dqkv
                             (always 0 today: per-column atomics from 148 CTAs contended and the extra registers
                             spilled in the epilogue -- measured 222 -> 282 us; the field stays for round 2) */
} missm_attn_args;

int missm_attention_fwd(const missm_attn_args* args, void* stream);
int missm_attention_bwd(missm_attn_args* args, void* stream);   /* writes args->colsum_done */

/* ---------------------------------------------------------------------------------------
 * LayerNorm over rows of the fp32 residual stream (nn.LayerNorm at modeling_image.py:120,132,
 * 139,149,649,660,514 and src/model/baseline.py:61).  Row r reads x[(row_index ? row_index[r]
 * : r) * ldx].  Optional add_rows: x <- x + add_rows[(r / add_div) % add_period] first (the
 * temporal embedding of modeling_image.py:110-114), written back to x_out.
 * ------------------------------------------------------------------------------------- */
int missm_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_index,
                        const float* add_rows, int32_t add_period, int32_t add_div, float* x_out,
                        const float* gamma, const float* beta, void* y, int64_t ldy, int32_t y_bf16,
                        float* mean, float* rstd, int32_t M, int32_t D, float eps, void* stream);
int missm_ln_bwd_num_partials(int32_t M);
/* dx[row_index? row_index[r] : r] = (dres ? dres : 0) + LN'(dy); partial = workspace of
 * missm_ln_bwd_num_partials(M) * 3 * D floats; dgamma, dbeta = [D]; dx_colsum (optional, [D]) =
 * sum over the written rows of dx (the bias gradient of the Linear that produced this stream). */
int missm_layernorm_bwd(const void* dy, int64_t lddy, int32_t dy_bf16, const float* x, int64_t ldx,
                        const int32_t* row_index, const float* mean, const float* rstd,
                        const float* gamma, const float* dres, float* dx, void* dx_bf16,
                        float* partial, float* dgamma, float* dbeta, float* dx_colsum, int32_t M,
                        int32_t D, void* stream);
int missm_reduce_partials(const float* partial, int32_t R, int64_t stride, float* out, int32_t n,
                          float scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * HBM-bound helpers
 * ------------------------------------------------------------------------------------- */
/* weights fp32 -> bf16 operand copies (dst may be wider: zero padded to cols_dst) */
int missm_cast_f32_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int32_t rows,
                        int32_t cols, int32_t cols_dst, void* stream);
/* bias gradients: out[n] = sum_m x[m, n]  (x bf16); partial = [missm_colsum_num_partials(M), N] */
int missm_colsum_num_partials(int32_t M);
int missm_colsum_bf16(const void* x, int64_t ldx, int32_t M, int32_t N, float* partial, float* out,
                      void* stream);
/* Conv2d(k = stride = ps, no bias) input patches (video/modeling_video.py:29-35,45):
 * pixels f32 [*, C, T, H, W] (T = 1 for images; the video reshape `b c t h w -> (b t) c h w` of
 * modeling_image.py:636-639 is folded in; sample b read from sample_index ? sample_index[b] : b)
 * -> bf16 [Bn * T * (H/ps) * (W/ps), Kpad], column (c*ps + i)*ps + j, zero padded */
int missm_patchify(const float* pixels, const int32_t* sample_index, void* patches, int32_t Bn,
                   int32_t C, int32_t T, int32_t H, int32_t W, int32_t ps, int32_t Kpad, void* stream);
/* The same convolution as an IMPLICIT GEMM on tcgen05 (csrc/patch_embed_tc.cu), forward only: reads the fp32 pixels
 * directly (same addressing as missm_patchify), multiplies by w_bf16 [D, Kpad] (Conv2d weight flattened to (c, i, j),
 * zero padded to Kpad), adds the position row and writes token rows -- what missm_patchify + missm_gemm_bf16 with
 * MISSM_EPI_PATCH do, without the [rows, Kpad] im2col matrix:
 *   tok[img * (P + 1) + 1 + patch, :] = patches(img, patch, :) . w^T + pos[1 + patch, :],  img = b * T + t, P = (H/ps)(W/ps)
 * (row img * (P + 1) is the CLS row: missm_cls_rows).  Returns 0 if launched, -1 if the shape is not handled
 * (ps != 14, Kpad > 640, D % 128 != 0): the caller then takes the explicit path. */
int missm_patch_embed_implicit(const float* pixels, const int32_t* sample_index, const void* w_bf16, const float* pos,
                               float* tok, int32_t Bn, int32_t C, int32_t T, int32_t H, int32_t W, int32_t ps,
                               int32_t Kpad, int32_t D, void* stream);
/* out[g] = sum of rows r of x f32 [M, D] with (r / div) % period == g  (temporal-embedding grad) */
int missm_colsum_grouped_f32(const float* x, int32_t M, int32_t D, int32_t period, int32_t div,
                             float* out, void* stream);
int missm_copy_f32(const float* src, float* dst, int64_t n, void* stream);
/* tok[b, 0, :] = class_embedding + position_embedding[0]   (modeling_video.py:48-50) */
int missm_cls_rows(const float* cls, const float* pos, float* tok, int32_t Bn, int32_t ntok, int32_t D,
                   void* stream);
/* dpos[t] = sum_b dtok[b, t]; dpatch_bf16[b*(ntok-1) + t-1] = dtok[b, t] for t >= 1 */
int missm_embed_bwd(const float* dtok, float* dpos, void* dpatch_bf16, int32_t Bn, int32_t ntok,
                    int32_t D, void* stream);
/* pooled.reshape(B, T, -1).mean(1)   (modeling_image.py:662) */
int missm_frame_mean(const float* in, void* out, int32_t out_bf16, int32_t Bn, int32_t T, int32_t D,
                     void* stream);
int missm_frame_mean_bwd(const float* dout, float* din, int32_t Bn, int32_t T, int32_t D, void* stream);
/* value / value.norm(p=2, dim=-1) * exp(logit_scale)   (languagebind/__init__.py:80-83) */
int missm_l2norm_scale_fwd(const float* x, float* y, float* inv_norm, float scale, int32_t Bn,
                           int32_t P, void* stream);
int missm_l2norm_scale_bwd(const float* dy, const float* x, const float* inv_norm, float scale,
                           void* dx, int32_t dx_bf16, int32_t Bn, int32_t P, void* stream);
/* transformers 4.3x CLIPTextEmbeddings (modeling_image.py:463,494) and the EOT pooling index
 * (modeling_image.py:519-522: argmax of int32-cast ids, first maximum) */
int missm_text_embed_fwd(const int64_t* ids, const int32_t* sample_index, const float* tok_emb,
                         const float* pos_emb, float* out, int32_t Bn, int32_t L, int32_t D,
                         void* stream);
int missm_text_embed_bwd(const int64_t* ids, const int32_t* sample_index, const float* dx,
                         float* dtok_emb, float* dpos, int32_t Bn, int32_t L, int32_t D, void* stream);
int missm_argmax_rows(const int64_t* ids, const int32_t* sample_index, int32_t* out_rows, int32_t Bn,
                      int32_t L, void* stream);

/* ---------------------------------------------------------------------------------------
 * Missing-modality mask compaction (mask of src/model/baseline.py:57).
 * For tower i: present_idx[i, 0:counts[i]] = ascending { b : missing_index[b] != codes[i] },
 * slot_of[i, b] = position of b in that list or -1.  codes_host is a HOST array.
 * ------------------------------------------------------------------------------------- */
#define MISSM_MAX_TOWERS 8
int missm_compact_mask(const int64_t* missing_index, int32_t Bn, const int32_t* codes_host,
                       int32_t n_towers, int32_t* present_idx, int32_t* slot_of, int32_t* counts,
                       void* stream);
/* dst[b] = slot_of[b] >= 0 ? src[slot_of[b]] : 0   (f32 rows of P) */
int missm_scatter_rows_zero(const float* src, const int32_t* slot_of, float* dst, int32_t Bn,
                            int32_t P, void* stream);
/* dst[r] = src[idx[r]]   (rows of row_bytes bytes, multiple of 16) */
int missm_gather_rows(const void* src, const int32_t* idx, void* dst, int32_t n_rows,
                      int64_t row_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Masked fusion of the default `sum` head, fp32 (modal_sum.forward, src/model/baseline.py:52-61):
 *   pre[b] = sum_m (missing_index[b] != codes[m]) * (emb_m[b] W_m^T + bias_m);  out = LayerNorm(pre)
 * emb_m f32 [B, P], weight_m f32 [Fd, P], bias_m f32 [Fd]; pre/out f32 [B, Fd]; mean/rstd f32 [B].
 * bwd: d_emb_m [B, P], d_weight_m [Fd, P], d_bias_m [Fd], d_gamma/d_beta [Fd];
 *      workspace = 3 * B * Fd floats.
 * ------------------------------------------------------------------------------------- */
typedef struct missm_fusion_sum_args {
  const float* emb[MISSM_MAX_TOWERS];
  const float* weight[MISSM_MAX_TOWERS];
  const float* bias[MISSM_MAX_TOWERS];
  float* d_emb[MISSM_MAX_TOWERS];
  float* d_weight[MISSM_MAX_TOWERS];
  float* d_bias[MISSM_MAX_TOWERS];
  int32_t codes[MISSM_MAX_TOWERS];
  const int64_t* missing_index;
  const float* gamma;
  const float* beta;
  float* pre;
  float* out;
  float* mean;
  float* rstd;
  int32_t n_modal, B, P, Fd;
  float eps;
} missm_fusion_sum_args;
int missm_fusion_sum_fwd(const missm_fusion_sum_args* args, void* stream);
int missm_fusion_sum_bwd(const missm_fusion_sum_args* args, const float* d_out, float* workspace,
                         float* d_gamma, float* d_beta, void* stream);

/* ---------------------------------------------------------------------------------------
 * fp32 VERIFICATION mode (host side: MISSM_PRECISION=fp32; csrc/fp32_mode.cu).  Same path, fp32-grade
 * arithmetic, to check embeddings / loss against the reference's fp32 PyTorch path at <= 1e-5.
 * An fp32 GEMM is ONE launch of missm_gemm_bf16 over 3-way bf16 splits of both operands laid out along the
 * contraction dimension (K' = 6 K; pieces x2,x3,x1,x2,x1,x1 against w2,w1,w3,w1,w2,w1, small products first): missm_expand6_bf16
 * builds such an operand from an fp32 matrix [rows, cols] (which = 0: A pattern, 1: B pattern;
 * stack_rows = 0: dst [rows, 6 * cols_pad] for K-major operands, zero padded; 1: dst [6 * rows, cols] for MN-major).
 * missm_attention_f32_*: the CLIPAttention chain in fp32 on the CUDA cores, same layout contract as
 * missm_attention_fwd/bwd but qkv / out / d_out / dqkv are float; any N <= 1024, masks supported.
 * ------------------------------------------------------------------------------------- */
int missm_expand6_bf16(const float* src, int64_t ld_src, int32_t rows, int32_t cols, void* dst, int64_t ld_dst,
                       int32_t cols_pad, int32_t which, int32_t stack_rows, void* stream);
int missm_patchify_f32(const float* pixels, const int32_t* sample_index, float* patches, int32_t Bn, int32_t C,
                       int32_t T, int32_t H, int32_t W, int32_t ps, int32_t Kpad,
                       void* stream);   /* patches f32 [rows, Kpad], pre-zeroed by the caller */
int missm_gelu_f32_fwd(const float* u, float* a, int64_t n, void* stream);
int missm_gelu_f32_bwd(const float* d_a, const float* u, float* d_u, int64_t n, void* stream);
int missm_attention_f32_fwd(const missm_attn_args* args, void* stream);
int missm_attention_f32_bwd(const missm_attn_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-tensor Adam (csrc/optim.cu) -- replaces torch.optim.Adam.step() as the reference drives it
 * (train_ddp.py:205 `optim.Adam(model.parameters(), lr, weight_decay)`, :254 `optimizer.step()`): L2 weight decay
 * folded into the gradient, no amsgrad, per-tensor step counts.  ONE launch updates every tensor of a parameter
 * group.  All table pointers are DEVICE arrays of n_tensors (chunk_*: n_chunks) entries built by the caller:
 * tensor t is cut into chunks of chunk_elems elements, chunk c covers elements
 * [chunk_offset[c], chunk_offset[c] + chunk_elems) of tensor chunk_tensor[c].  grads[t] == NULL skips tensor t
 * (torch skips parameters without a gradient).  step_size[t] = lr / (1 - beta1^step_t), bc2_sqrt[t] =
 * sqrt(1 - beta2^step_t), computed by the caller in double.  bf16_out (may be NULL, entries may be NULL): also
 * write the bf16 GEMM-operand copy of the updated parameter.  zero_grads: also zero the gradients
 * (zero_grad(set_to_none=False) fused).  Everything fp32, contiguous; p / g / m / v 16-byte aligned tensors take
 * the 128-bit path.  HBM-bound: 28 B per element (+2 bf16_out, +4 zero_grads).
 * ------------------------------------------------------------------------------------- */
typedef struct missm_adam_args {
  void* const* params;
  void* const* grads;
  void* const* exp_avg;
  void* const* exp_avg_sq;
  void* const* bf16_out;
  const int64_t* numel;
  const float* step_size;
  const float* bc2_sqrt;
  const int32_t* chunk_tensor;
  const int64_t* chunk_offset;
  int64_t chunk_elems;
  int32_t n_tensors, n_chunks;
  double beta1, beta2;
  float eps, weight_decay;
  int32_t zero_grads;
} missm_adam_args;
int missm_adam_multi(const missm_adam_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * GPU input pipeline for the image-shaped modalities (csrc/preprocess.cu) -- replaces the per-sample torchvision
 * chain of the reference's processors: ToTensor -> Resize(S, BICUBIC) -> CenterCrop(S) -> Normalize
 * (languagebind/image/processing_image.py:20-29, thermal/processing_thermal.py:15-25) and, for depth,
 * DepthNorm in front of it (depth/processing_depth.py:21-57).  One launch per decoded image:
 *   v   = clamp(src / pre_div, clip_lo, clip_hi) / post_div        (ToTensor: pre_div 255, no clip, post_div 1;
 *                                                                   DepthNorm: 1000, [0.01, max_depth], max_depth)
 *   r   = bicubic resample of v so that the SHORTER side becomes S (the other int(S * long / short)),
 *         align_corners = False; antialias = 0: 4-tap cubic convolution, A = -0.75, clamped border indices
 *         (torchvision <= 0.16 on tensors); antialias = 1: area-scaled filter, A = -0.5, support 2 * scale,
 *         truncated + renormalised at the border (torchvision >= 0.17 default)
 *   dst = (center S x S crop of r - mean[c]) / std[c]              fp32 [3, S, S], device
 * src: device, uint8 [H, W, 3] (src_f32 = 0) or float [H, W] replicated to 3 channels (src_f32 = 1).
 * ------------------------------------------------------------------------------------- */
typedef struct missm_preproc_args {
  const void* src;
  float* dst;
  int32_t src_f32, H, W, S, antialias;
  float pre_div, clip_lo, clip_hi, post_div;
  float mean[3];
  float std_[3];
} missm_preproc_args;
int missm_image_preprocess(const missm_preproc_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Residual-block drivers (csrc/blocks.cu): ONE call issues every kernel of one residual block of
 * CLIPEncoderLayer.forward (languagebind/image/modeling_image.py:105-127 temporal attention, :129-134 temporal MLP,
 * :137-146 spatial attention, :148-151 MLP; the autograd twins run at train_ddp.py:253) into caller-provided
 * memory, so the host pays one binding call per block instead of one per kernel (round 1: 1 936 binding calls and
 * 98.6 ms of host time per 101.6 ms step).  Nothing allocates; `saved` lives from forward to backward, `scratch` only
 * during the backward call (stream-ordered: the caller may recycle it as soon as the call returns if the next user
 * is on the same stream).  missm_*_block_sizes fills {saved bytes, scratch bytes, grads floats}.
 *
 * attention block:  out = x' + OutProj(Attention(QKV(LN(x'))))      x' = x + add_rows[(r / add_div) % add_period]
 *   w_qkv bf16 [3D, ldw_qkv] = q | k | v rows, b_qkv f32 [3D]; w_o bf16 [D, ldw_o], b_o f32 [D]; the q block of the
 *   projection is scaled by head_dim^-0.5 in the GEMM epilogue; sequence layout / masks as missm_attn_args.
 *   LoRA (lora_r > 0; peft Linear y = W x + b + s B(A x), reference convert_to_lora, modeling_image.py:775-793), with
 *   R3 = pad64(3 r), R1 = pad64(r) (rank groups padded with zeros to one 64-element K block of the GEMM tiles, so
 *   that the contraction dimension D + R stays a multiple of the tile depth and no TMA box is out of bounds): w_qkv = [W | s B_cat] (ldw_qkv >= D + R3), wb_qkv bf16 [3D + R3, D] = [W ; A_cat]
 *   row-stacked; w_o = [W_o | s B_o] (ldw_o >= D + R1), wb_o bf16 [D + R1, D] = [W_o ; A_o].
 *   grads (floats, in this order): d_ln_w [D], d_ln_b [D], dx_colsum [D], d_w_qkv [3D, D], d_b_qkv [3D],
 *   d_w_o [D, D], d_b_o [D], d_add_rows [add_period, D] (if add_rows), then with LoRA d_A_cat [R3, D],
 *   d_sB_cat [3D, R3], d_A_o [R1, D], d_sB_o [D, R1] (gradients w.r.t. the PACKED operands: the caller slices the
 *   q / k / v diagonal blocks and multiplies the B parts by s).  wgrad = 0 (frozen encoder): d_w_*, d_b_* are not
 *   computed.  dx_colsum = column sums of dx = bias gradient of the Linear that produced x.
 *   Backward inputs: d_out f32 [M, D] (required), d_out_bf16 (optional bf16 copy [M, D]; else made here),
 *   d_out_colsum_given: 1 = the caller already owns d_b_o (the dx_colsum of the block that consumed `out`), so it is
 *   not recomputed.  Outputs: dx f32 [M, D], dx_bf16 [M, D].
 * MLP block:  out = x + fc2(quick_gelu(fc1(LN(x))));  w1 bf16 [F, D], w2 bf16 [D, F];
 *   grads: d_ln_w [D], d_ln_b [D], dx_colsum [D], d_w1 [F, D], d_b1 [F], d_w2 [D, F], d_b2 [D].
 * ------------------------------------------------------------------------------------- */
typedef struct missm_attn_block_args {
  int32_t M, D, H;
  float eps;
  int64_t seq_outer, seq_inner, tok_stride;
  int32_t N, n_seq, s_in, causal, mask_div;
  const int64_t* key_mask;
  const int32_t* mask_rows;
  const float* add_rows;
  int32_t add_period, add_div;
  const float* ln_w;
  const float* ln_b;
  const void* w_qkv;
  const float* b_qkv;
  const void* w_o;
  const float* b_o;
  int32_t ldw_qkv, ldw_o;
  int32_t lora_r;
  const void* wb_qkv;
  const void* wb_o;
  const float* x;
  float* out;
  void* saved;
  /* backward */
  const float* d_out;
  const void* d_out_bf16;
  int32_t d_out_colsum_given;
  int32_t wgrad;
  float* dx;
  void* dx_bf16;
  float* grads;
  void* scratch;
} missm_attn_block_args;
int missm_attn_block_sizes(const missm_attn_block_args* args, int64_t sizes[3]);
int missm_attn_block_fwd(const missm_attn_block_args* args, void* stream);
int missm_attn_block_bwd(const missm_attn_block_args* args, void* stream);

typedef struct missm_mlp_block_args {
  int32_t M, D, F;
  float eps;
  const float* ln_w;
  const float* ln_b;
  const void* w1;
  const float* b1;
  const void* w2;
  const float* b2;
  const float* x;
  float* out;
  void* saved;
  /* backward */
  const float* d_out;
  const void* d_out_bf16;
  int32_t d_out_colsum_given;
  int32_t wgrad;
  float* dx;
  void* dx_bf16;
  float* grads;
  void* scratch;
} missm_mlp_block_args;
int missm_mlp_block_sizes(const missm_mlp_block_args* args, int64_t sizes[3]);
int missm_mlp_block_fwd(const missm_mlp_block_args* args, void* stream);
int missm_mlp_block_bwd(const missm_mlp_block_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * GPU input pipeline, video and audio (csrc/preprocess_av.cu).
 * missm_video_preprocess: the transform chain of languagebind/video/processing_video.py:25-66 on the decoded frames of
 *   one clip: x / 255 -> Normalize(mean, std) -> ShortSideScale(S) (bilinear, align_corners = False, no antialias)
 *   -> CenterCrop(S) -> optional horizontal flip.  src: device uint8 [T, H, W, 3]; dst: device f32 [3, T, S, S].
 * missm_audio_fbank: torchaudio.compliance.kaldi.fbank as audio/processing_audio.py:96-110 calls it (16 kHz: 400-sample
 *   hanning frames every 160 samples, DC removal, pre-emphasis 0.97, 512-point power spectrum, n_mel triangular mel
 *   filters given by the caller as mel_weights f32 [n_mel, 257], log) followed by waveform2melspec's tail (:53-94):
 *   out[c, m, t] = (mel[(offsets[c] + t) % n_frames, m] - mean) / (2 std), f32 [3, n_mel, target].  wave: channel 0
 *   (n_samples); wave_all / n_total: the whole loaded tensor whose mean is subtracted first (`audio_data -=
 *   audio_data.mean()`); mel: workspace f32 [missm_fbank_num_frames(n_samples), n_mel]; wave_sum: workspace, 1 double.
 * ------------------------------------------------------------------------------------- */
typedef struct missm_video_args {
  const void* src;
  float* dst;
  int32_t T, H, W, S, hflip;
  float mean[3];
  float std_[3];
} missm_video_args;
int missm_video_preprocess(const missm_video_args* args, void* stream);

typedef struct missm_fbank_args {
  const float* wave;
  const float* wave_all;
  int64_t n_samples, n_total;
  const float* mel_weights;
  float* mel;
  double* wave_sum;
  float* out;
  int32_t n_mel, target;
  int32_t offsets[3];
  float mean, std_;
} missm_fbank_args;
int missm_fbank_num_frames(int64_t n_samples);
int missm_audio_fbank(const missm_fbank_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Evaluation metrics accumulated on the device (csrc/metrics.cu) -- replaces the per-batch host round trips of the
 * reference's evaluate() (train_ddp.py:88-133, test.py:21-66: `loss.item()`, argmax / softmax `.cpu().numpy()` every
 * batch).  One launch per batch: confusion[label, argmax] += 1 (int64 [C, C], first maximum as torch.argmax),
 * loss_sum += mean CrossEntropy of the batch (double), probs[b, :] = softmax(logits[b, :]) (for roc_auc_score),
 * n_seen += B.  Accuracy and macro-F1 follow from the confusion matrix at the epoch's single read-out
 * (missm_b200/metrics.py).  logits f32 [B, C] contiguous, labels int64 [B]; accumulators are caller-zeroed.
 * ------------------------------------------------------------------------------------- */
int missm_eval_accumulate(const float* logits, const int64_t* labels, int32_t B, int32_t C, float* probs,
                          int64_t* confusion, double* loss_sum, int64_t* n_seen, void* stream);

/* ---------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): the number of kernels this library has launched since the last reset, and CUDA
 * events around every missm_gemm_bf16 launch (on the launching stream) while the profile is on.
 * missm_gemm_profile(1) starts (and clears), missm_gemm_profile(0) stops; missm_gemm_profile_read synchronises the
 * recorded events and returns the summed milliseconds, 2*M*N*K flops and launch count.
 * ------------------------------------------------------------------------------------- */
int64_t missm_launch_count(int32_t reset);
/* debugging aid (MISSM_DEBUG_EVENTS=1 records a CUDA event after every kernel-level call of the block drivers):
 * prints, per stream, the first recorded call that has not finished -- what a stalled GPU is stuck on */
int missm_debug_dump(void);
int missm_debug_crumb(const char* label_static, void* stream);   /* record one more event (label must stay alive) */
int missm_gemm_profile(int32_t on);
int missm_gemm_profile_read(double* ms, double* flop, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* MISSM_B200_H_ */
