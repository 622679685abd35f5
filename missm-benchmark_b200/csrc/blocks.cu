// Residual-block drivers: one C-ABI call issues every kernel of one residual block of CLIPEncoderLayer
// (reference: languagebind/image/modeling_image.py:105-151; CLIPAttention / CLIPMLP = transformers 4.3x) into
// caller-provided memory.  No kernels live here -- this file is the host-side schedule of a block over the
// library's own entry points (tcgen05 GEMMs with fused epilogues, fused attention, LayerNorm, reductions).
//
// Why it exists: driven kernel by kernel from Python the training step was HOST-bound (round 1: 1 936 binding calls,
// each with its own output allocation and argument struct, 98.6 ms of issue time per 101.6 ms step).  A block is
// 5 (forward) / 11-13 (backward) launches; issuing them from C costs ~4 us each instead of ~50.
//
// Memory layout of `saved` / `scratch` / `grads`: see the *_layout structs below; every sub-buffer is 256-byte
// aligned (TMA base addresses need 16).
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/missm_b200.h"
#include "missm_common.cuh"

using namespace missm;

namespace {

constexpr int64_t kAlign = 256;
inline int64_t up(int64_t v) { return (v + kAlign - 1) / kAlign * kAlign; }
inline int pad8(int n) { return (n + 63) / 64 * 64; }   // LoRA rank groups: padded to one 64-element K block

struct Carver {
  char* base;
  int64_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(int64_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += up(count * static_cast<int64_t>(sizeof(T)));
    return p;
  }
};

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ attention
struct AttnSaved {
  float* x_res;   // [M, D]   only with add_rows
  float* mean;    // [M]
  float* rstd;    // [M]
  bf16* hcat;     // [M, D + R3]      LN(x') | LN(x') A_cat^T
  bf16* qkvcat;   // [M, 3D + R3]     q | k | v (| pitch shared with [dqkv | dT] in the backward)
  bf16* attncat;  // [M, D + R1]      attention output | attn A_o^T
  float* lse;     // [n_seq, H, N]
  int64_t bytes;
};
struct AttnScratch {
  bf16* dycat;     // [M, D + R1]   bf16(d_out) | d_out (sB_o)
  bf16* d_attncat; // [M, D + R1]
  bf16* dqkvcat;   // [M, 3D + R3]
  bf16* d_h;       // [M, D]
  float* delta;    // [n_seq, H, N]
  float* ln_part;  // [ln partials, 3, D]
  float* cs_part;  // [colsum partials, 3D]
  int64_t bytes;
};
struct AttnGrads {
  float *ln_w, *ln_b, *dx_colsum, *w_qkv, *b_qkv, *w_o, *b_o, *add_rows, *a_cat, *sb_cat, *a_o, *sb_o;
  int64_t floats;
};

inline int r3_of(const missm_attn_block_args* a) { return a->lora_r > 0 ? pad8(3 * a->lora_r) : 0; }
inline int r1_of(const missm_attn_block_args* a) { return a->lora_r > 0 ? pad8(a->lora_r) : 0; }

AttnSaved attn_saved(const missm_attn_block_args* a, void* base) {
  const int64_t M = a->M, D = a->D;
  Carver c(base);
  AttnSaved s;
  s.x_res = a->add_rows ? c.take<float>(M * D) : nullptr;
  s.mean = c.take<float>(M);
  s.rstd = c.take<float>(M);
  s.hcat = c.take<bf16>(M * (D + r3_of(a)));
  s.qkvcat = c.take<bf16>(M * (3 * D + r3_of(a)));
  s.attncat = c.take<bf16>(M * (D + r1_of(a)));
  s.lse = c.take<float>(static_cast<int64_t>(a->n_seq) * a->H * a->N);
  s.bytes = c.off;
  return s;
}
AttnScratch attn_scratch(const missm_attn_block_args* a, void* base) {
  const int64_t M = a->M, D = a->D;
  Carver c(base);
  AttnScratch s;
  s.dycat = c.take<bf16>(M * (D + r1_of(a)));
  s.d_attncat = c.take<bf16>(M * (D + r1_of(a)));
  s.dqkvcat = c.take<bf16>(M * (3 * D + r3_of(a)));
  s.d_h = c.take<bf16>(M * D);
  s.delta = c.take<float>(static_cast<int64_t>(a->n_seq) * a->H * a->N);
  s.ln_part = c.take<float>(static_cast<int64_t>(missm_ln_bwd_num_partials(a->M)) * 3 * D);
  s.cs_part = c.take<float>(static_cast<int64_t>(missm_colsum_num_partials(a->M)) * 3 * D);
  s.bytes = c.off;
  return s;
}
AttnGrads attn_grads(const missm_attn_block_args* a, float* base) {
  const int64_t D = a->D;
  AttnGrads g;
  int64_t o = 0;
  auto take = [&](int64_t n) { float* p = base ? base + o : nullptr; o += n; return p; };
  g.ln_w = take(D), g.ln_b = take(D), g.dx_colsum = take(D);
  g.w_qkv = take(3 * D * D), g.b_qkv = take(3 * D), g.w_o = take(D * D), g.b_o = take(D);
  g.add_rows = a->add_rows ? take(static_cast<int64_t>(a->add_period) * D) : nullptr;
  g.a_cat = g.sb_cat = g.a_o = g.sb_o = nullptr;
  if (a->lora_r > 0) {
    g.a_cat = take(r3_of(a) * D), g.sb_cat = take(3 * D * r3_of(a));
    g.a_o = take(r1_of(a) * D), g.sb_o = take(D * r1_of(a));
  }
  g.floats = o;
  return g;
}

int check_attn(const missm_attn_block_args* a) {
  MISSM_REQUIRE(a != nullptr, "attn_block: null args");
  MISSM_REQUIRE(a->M > 0 && a->D > 0 && a->H > 0 && a->D == a->H * 64, "attn_block: M=%d D=%d H=%d (head_dim must be 64)",
                a->M, a->D, a->H);
  MISSM_REQUIRE(a->D % 128 == 0, "attn_block: D=%d must be a multiple of 128", a->D);
  MISSM_REQUIRE(a->lora_r >= 0 && (a->lora_r == 0 || (a->wb_qkv && a->wb_o)), "attn_block: LoRA needs wb_qkv / wb_o");
  MISSM_REQUIRE(a->ldw_qkv >= a->D + r3_of(a) && a->ldw_o >= a->D + r1_of(a) && a->ldw_qkv % 8 == 0 && a->ldw_o % 8 == 0,
                "attn_block: weight pitches %d / %d", a->ldw_qkv, a->ldw_o);
  MISSM_REQUIRE(a->add_rows == nullptr || (a->add_period > 0 && a->add_div > 0), "attn_block: add_rows needs period / div");
  return 0;
}

int gemm(const void* A, int lda, bool a_mn, const void* B, int ldb, bool b_mn, void* C, int ldc, bool c_f32, int M, int N,
         int K, void* stream, int epi = MISSM_EPI_LINEAR, const float* bias = nullptr, const void* aux_in = nullptr,
         int ld_aux_in = 0, void* aux_out = nullptr, int ld_aux_out = 0, int scale_cols = 0, float col_scale = 1.f,
         float* colsum_part = nullptr) {
  missm_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = A, g.B = B, g.C = C, g.bias = bias, g.aux_in = aux_in, g.aux_out = aux_out;
  g.M = M, g.N = N, g.K = K, g.lda = lda, g.ldb = ldb, g.ldc = ldc, g.ld_aux_in = ld_aux_in, g.ld_aux_out = ld_aux_out;
  g.a_mn = a_mn, g.b_mn = b_mn, g.epilogue = epi, g.out_f32 = c_f32;
  g.scale_cols = scale_cols, g.col_scale = col_scale;
  g.colsum_part = colsum_part;
  return missm_gemm_bf16(&g, stream);
}

void fill_attn(missm_attn_args& t, const missm_attn_block_args* a, const bf16* qkv, int ld_qkv, bf16* out, int ld_o, float* lse) {
  memset(&t, 0, sizeof(t));
  t.qkv = qkv, t.out = out, t.lse = lse;
  t.key_mask = a->key_mask, t.mask_rows = a->mask_rows;
  t.ld_qkv = ld_qkv, t.ld_o = ld_o;
  t.seq_outer = a->seq_outer, t.seq_inner = a->seq_inner, t.tok_stride = a->tok_stride;
  t.D = a->D, t.H = a->H, t.N = a->N, t.head_dim = 64;
  t.n_seq = a->n_seq, t.s_in = a->s_in, t.causal = a->causal, t.mask_div = a->mask_div;
}

// ---- debugging aid (MISSM_DEBUG_EVENTS=1): a CUDA event after every kernel-level call of a driver, so that a
// stalled GPU can be asked which call never finished (missm_debug_dump; used by scratch/hang_watch.py)
struct Crumb { cudaEvent_t ev; const char* what; void* stream; int line; };
std::mutex g_crumb_mu;
std::vector<Crumb> g_crumbs;
const bool g_debug_events = getenv("MISSM_DEBUG_EVENTS") != nullptr;
void crumb(const char* what, int line, void* stream) {
  Crumb c{nullptr, what, stream, line};
  cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming);
  cudaEventRecord(c.ev, static_cast<cudaStream_t>(stream));
  std::lock_guard<std::mutex> lk(g_crumb_mu);
  g_crumbs.push_back(c);
}

#define RC(expr)                                            \
  do {                                                      \
    if (int _rc = (expr)) return _rc;                       \
    if (g_debug_events) crumb(#expr, __LINE__, stream);     \
  } while (0)

}  // namespace

extern "C" int missm_attn_block_sizes(const missm_attn_block_args* a, int64_t sizes[3]) {
  if (int rc = check_attn(a)) return rc;
  sizes[0] = attn_saved(a, nullptr).bytes;
  sizes[1] = attn_scratch(a, nullptr).bytes;
  sizes[2] = attn_grads(a, nullptr).floats;
  return 0;
}

extern "C" int missm_attn_block_fwd(const missm_attn_block_args* a, void* stream) {
  RC(check_attn(a));
  MISSM_REQUIRE(a->x && a->out && a->saved && a->ln_w && a->ln_b && a->w_qkv && a->b_qkv && a->w_o && a->b_o,
                "attn_block_fwd: null pointer");
  const int M = a->M, D = a->D, R3 = r3_of(a), R1 = r1_of(a);
  const AttnSaved s = attn_saved(a, a->saved);
  const float* x_res = a->add_rows ? s.x_res : a->x;
  const int ld_h = D + R3, ld_qkv = 3 * D + R3, ld_at = D + R1;
  RC(missm_layernorm_fwd(a->x, D, nullptr, a->add_rows, a->add_period, a->add_div, s.x_res, a->ln_w, a->ln_b, s.hcat, ld_h, 1,
                         s.mean, s.rstd, M, D, a->eps, stream));
  if (R3)   // T = h A_cat^T, written next to h
    RC(gemm(s.hcat, ld_h, false, static_cast<const bf16*>(a->wb_qkv) + 3L * D * D, D, false, s.hcat + D, ld_h, false, M, R3, D, stream));
  RC(gemm(s.hcat, ld_h, false, a->w_qkv, a->ldw_qkv, false, s.qkvcat, ld_qkv, false, M, 3 * D, D + R3, stream, MISSM_EPI_LINEAR,
          a->b_qkv, nullptr, 0, nullptr, 0, D, 0.125f));
  missm_attn_args t;
  fill_attn(t, a, s.qkvcat, ld_qkv, s.attncat, ld_at, s.lse);
  RC(missm_attention_fwd(&t, stream));
  if (R1)
    RC(gemm(s.attncat, ld_at, false, static_cast<const bf16*>(a->wb_o) + static_cast<long>(D) * D, D, false, s.attncat + D, ld_at,
            false, M, R1, D, stream));
  RC(gemm(s.attncat, ld_at, false, a->w_o, a->ldw_o, false, a->out, D, true, M, D, D + R1, stream, MISSM_EPI_RESID, a->b_o, x_res, D));
  return 0;
}

extern "C" int missm_attn_block_bwd(const missm_attn_block_args* a, void* stream) {
  RC(check_attn(a));
  MISSM_REQUIRE(a->saved && a->scratch && a->grads && a->d_out && a->dx && a->dx_bf16 && a->w_qkv && a->w_o && a->ln_w,
                "attn_block_bwd: null pointer");
  const int M = a->M, D = a->D, R3 = r3_of(a), R1 = r1_of(a);
  const AttnSaved s = attn_saved(a, a->saved);
  const AttnScratch w = attn_scratch(a, a->scratch);
  const AttnGrads g = attn_grads(a, a->grads);
  const float* x_res = a->add_rows ? s.x_res : a->x;
  MISSM_REQUIRE(x_res != nullptr, "attn_block_bwd: x is needed (LayerNorm backward)");
  const int ld_h = D + R3, ld_qkv = 3 * D + R3, ld_at = D + R1;
  const bool wg = a->wgrad != 0;

  // ---- bf16 copy of d_out (+ its column sums = d_b_o) unless the consumer block handed them over
  const bf16* dy = static_cast<const bf16*>(a->d_out_bf16);
  int ld_dy = D;
  if (dy == nullptr || R1) {
    // LoRA needs dY inside the pitched [dY | dT_o] buffer: one cast pass from the fp32 gradient either way
    RC(missm_cast_f32_bf16(a->d_out, D, w.dycat, ld_at, M, D, D, stream));
    dy = w.dycat, ld_dy = ld_at;
  }
  if (wg && !a->d_out_colsum_given) RC(missm_colsum_bf16(dy, ld_dy, M, D, w.cs_part, g.b_o, stream));

  // ---- out_proj
  if (R1) {
    // dT_o = dY (sB_o); d_attn = [dY | dT_o] [W_o ; A_o]; adapter wgrads
    RC(gemm(dy, ld_dy, false, static_cast<const bf16*>(a->w_o) + D, a->ldw_o, true, w.dycat + D, ld_at, false, M, R1, D, stream));
    RC(gemm(w.dycat, ld_at, false, a->wb_o, D, true, w.d_attncat, ld_at, false, M, D, D + R1, stream));
    RC(gemm(w.dycat + D, ld_at, true, s.attncat, ld_at, true, g.a_o, D, true, R1, D, M, stream));          // dT_o^T attn
    RC(gemm(dy, ld_dy, true, s.attncat + D, ld_at, true, g.sb_o, R1, true, D, R1, M, stream));             // dY^T T_o
  } else {
    RC(gemm(dy, ld_dy, false, a->w_o, a->ldw_o, true, w.d_attncat, ld_at, false, M, D, D, stream));        // dY W_o
  }
  if (wg) RC(gemm(dy, ld_dy, true, s.attncat, ld_at, true, g.w_o, D, true, D, D, M, stream));              // dY^T attn

  // ---- attention core
  missm_attn_args t;
  fill_attn(t, a, s.qkvcat, ld_qkv, s.attncat, ld_at, s.lse);
  t.d_out = w.d_attncat, t.delta = w.delta, t.dqkv = w.dqkvcat, t.q_scale = 0.125f;
  RC(missm_attention_bwd(&t, stream));
  if (wg && !t.colsum_done) RC(missm_colsum_bf16(w.dqkvcat, ld_qkv, M, 3 * D, w.cs_part, g.b_qkv, stream));

  // ---- q / k / v projections
  if (R3) {
    RC(gemm(w.dqkvcat, ld_qkv, false, static_cast<const bf16*>(a->w_qkv) + D, a->ldw_qkv, true, w.dqkvcat + 3 * D, ld_qkv, false, M,
            R3, 3 * D, stream));                                                                           // dT = dqkv (sB_cat)
    RC(gemm(w.dqkvcat, ld_qkv, false, a->wb_qkv, D, true, w.d_h, D, false, M, D, 3 * D + R3, stream));     // dqkv W + dT A_cat
    RC(gemm(w.dqkvcat + 3 * D, ld_qkv, true, s.hcat, ld_h, true, g.a_cat, D, true, R3, D, M, stream));     // dT^T h
    RC(gemm(w.dqkvcat, ld_qkv, true, s.hcat + D, ld_h, true, g.sb_cat, R3, true, 3 * D, R3, M, stream));   // dqkv^T T
  } else {
    RC(gemm(w.dqkvcat, ld_qkv, false, a->w_qkv, a->ldw_qkv, true, w.d_h, D, false, M, D, 3 * D, stream));  // dqkv W
  }
  if (wg) RC(gemm(w.dqkvcat, ld_qkv, true, s.hcat, ld_h, true, g.w_qkv, D, true, 3 * D, D, M, stream));    // dqkv^T h

  // ---- LayerNorm backward + residual-stream gradient (fp32 + bf16 copy + column sums)
  RC(missm_layernorm_bwd(w.d_h, D, 1, x_res, D, nullptr, s.mean, s.rstd, a->ln_w, a->d_out, a->dx, a->dx_bf16, w.ln_part,
                         g.ln_w, g.ln_b, g.dx_colsum, M, D, stream));
  if (a->add_rows) RC(missm_colsum_grouped_f32(a->dx, M, D, a->add_period, a->add_div, g.add_rows, stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------------ MLP
namespace {
struct MlpSaved {
  float *mean, *rstd;
  bf16 *h, *u, *act;
  int64_t bytes;
};
struct MlpScratch {
  bf16 *dy, *d_u, *d_h;
  float *ln_part, *cs_part;
  int64_t bytes;
};
struct MlpGrads {
  float *ln_w, *ln_b, *dx_colsum, *w1, *b1, *w2, *b2;
  int64_t floats;
};
MlpSaved mlp_saved(const missm_mlp_block_args* a, void* base) {
  const int64_t M = a->M, D = a->D, F = a->F;
  Carver c(base);
  MlpSaved s;
  s.mean = c.take<float>(M), s.rstd = c.take<float>(M);
  s.h = c.take<bf16>(M * D), s.u = c.take<bf16>(M * F), s.act = c.take<bf16>(M * F);
  s.bytes = c.off;
  return s;
}
MlpScratch mlp_scratch(const missm_mlp_block_args* a, void* base) {
  const int64_t M = a->M, D = a->D, F = a->F;
  Carver c(base);
  MlpScratch s;
  s.dy = c.take<bf16>(M * D), s.d_u = c.take<bf16>(M * F), s.d_h = c.take<bf16>(M * D);
  s.ln_part = c.take<float>(static_cast<int64_t>(missm_ln_bwd_num_partials(a->M)) * 3 * D);
  const int cs_rows = missm_colsum_num_partials(a->M) > missm_gemm_colsum_rows(a->M) ? missm_colsum_num_partials(a->M)
                                                                                      : missm_gemm_colsum_rows(a->M);
  s.cs_part = c.take<float>(static_cast<int64_t>(cs_rows) * (F > D ? F : D));
  s.bytes = c.off;
  return s;
}
MlpGrads mlp_grads(const missm_mlp_block_args* a, float* base) {
  const int64_t D = a->D, F = a->F;
  MlpGrads g;
  int64_t o = 0;
  auto take = [&](int64_t n) { float* p = base ? base + o : nullptr; o += n; return p; };
  g.ln_w = take(D), g.ln_b = take(D), g.dx_colsum = take(D);
  g.w1 = take(F * D), g.b1 = take(F), g.w2 = take(D * F), g.b2 = take(D);
  g.floats = o;
  return g;
}
int check_mlp(const missm_mlp_block_args* a) {
  MISSM_REQUIRE(a != nullptr, "mlp_block: null args");
  MISSM_REQUIRE(a->M > 0 && a->D > 0 && a->F > 0 && a->D % 128 == 0 && a->F % 8 == 0, "mlp_block: M=%d D=%d F=%d", a->M, a->D, a->F);
  return 0;
}
}  // namespace

extern "C" int missm_mlp_block_sizes(const missm_mlp_block_args* a, int64_t sizes[3]) {
  if (int rc = check_mlp(a)) return rc;
  sizes[0] = mlp_saved(a, nullptr).bytes;
  sizes[1] = mlp_scratch(a, nullptr).bytes;
  sizes[2] = mlp_grads(a, nullptr).floats;
  return 0;
}

extern "C" int missm_mlp_block_fwd(const missm_mlp_block_args* a, void* stream) {
  RC(check_mlp(a));
  MISSM_REQUIRE(a->x && a->out && a->saved && a->ln_w && a->ln_b && a->w1 && a->b1 && a->w2 && a->b2, "mlp_block_fwd: null pointer");
  const int M = a->M, D = a->D, F = a->F;
  const MlpSaved s = mlp_saved(a, a->saved);
  RC(missm_layernorm_fwd(a->x, D, nullptr, nullptr, 0, 0, nullptr, a->ln_w, a->ln_b, s.h, D, 1, s.mean, s.rstd, M, D, a->eps, stream));
  RC(gemm(s.h, D, false, a->w1, D, false, s.act, F, false, M, F, D, stream, MISSM_EPI_GELU, a->b1, nullptr, 0, s.u, F));
  RC(gemm(s.act, F, false, a->w2, F, false, a->out, D, true, M, D, F, stream, MISSM_EPI_RESID, a->b2, a->x, D));
  return 0;
}

extern "C" int missm_mlp_block_bwd(const missm_mlp_block_args* a, void* stream) {
  RC(check_mlp(a));
  MISSM_REQUIRE(a->saved && a->scratch && a->grads && a->d_out && a->dx && a->dx_bf16 && a->x && a->w1 && a->w2 && a->ln_w,
                "mlp_block_bwd: null pointer");
  const int M = a->M, D = a->D, F = a->F;
  const MlpSaved s = mlp_saved(a, a->saved);
  const MlpScratch w = mlp_scratch(a, a->scratch);
  const MlpGrads g = mlp_grads(a, a->grads);
  const bool wg = a->wgrad != 0;
  const bf16* dy = static_cast<const bf16*>(a->d_out_bf16);
  if (dy == nullptr) {
    RC(missm_cast_f32_bf16(a->d_out, D, w.dy, D, M, D, D, stream));
    dy = w.dy;
  }
  if (wg && !a->d_out_colsum_given) RC(missm_colsum_bf16(dy, D, M, D, w.cs_part, g.b2, stream));
  if (wg) RC(gemm(dy, D, true, s.act, F, true, g.w2, F, true, D, F, M, stream));                            // dY^T a
  // (dY W2) gelu'(u); its epilogue also leaves the column sums of every 32 rows of d_u (= the fc1 bias gradient after one
  // small reduction): no second pass over the [M, F] gradient
  static const bool fused_cs = getenv("MISSM_COLSUM_EPILOGUE") == nullptr || atoi(getenv("MISSM_COLSUM_EPILOGUE")) != 0;   // A/B
  RC(gemm(dy, D, false, a->w2, F, true, w.d_u, F, false, M, F, D, stream, MISSM_EPI_DGELU, nullptr, s.u, F, nullptr, 0, 0, 1.f,
          wg && fused_cs ? w.cs_part : nullptr));
  if (wg) {
    if (fused_cs)
      RC(missm_reduce_partials(w.cs_part, missm_gemm_colsum_rows(M), F, g.b1, F, 1.0f, stream));
    else
      RC(missm_colsum_bf16(w.d_u, F, M, F, w.cs_part, g.b1, stream));
    RC(gemm(w.d_u, F, true, s.h, D, true, g.w1, D, true, F, D, M, stream));                                 // d_u^T h
  }
  RC(gemm(w.d_u, F, false, a->w1, D, true, w.d_h, D, false, M, D, F, stream));                              // d_u W1
  RC(missm_layernorm_bwd(w.d_h, D, 1, a->x, D, nullptr, s.mean, s.rstd, a->ln_w, a->d_out, a->dx, a->dx_bf16, w.ln_part, g.ln_w,
                         g.ln_b, g.dx_colsum, M, D, stream));
  return 0;
}

extern "C" int missm_debug_crumb(const char* what_static, void* stream) {
  if (g_debug_events) crumb(what_static, 0, stream);
  return 0;
}

// debugging aid: for every stream, the first recorded call whose event has not completed (see `crumb` above)
extern "C" int missm_debug_dump(void) {
  std::lock_guard<std::mutex> lk(g_crumb_mu);
  std::vector<void*> seen;
  fprintf(stderr, "missm_debug_dump: %zu recorded calls\n", g_crumbs.size());
  for (size_t i = 0; i < g_crumbs.size(); ++i) {
    const Crumb& c = g_crumbs[i];
    bool done = false;
    for (void* s : seen) done = done || s == c.stream;
    if (done) continue;
    if (cudaEventQuery(c.ev) == cudaSuccess) continue;
    seen.push_back(c.stream);
    fprintf(stderr, "  stream %p: first unfinished call #%zu (blocks.cu:%d) %.160s\n", c.stream, i, c.line, c.what);
    for (size_t j = i; j-- > 0;)
      if (g_crumbs[j].stream == c.stream) {
        fprintf(stderr, "      previous call on that stream (finished): blocks.cu:%d %.120s\n", g_crumbs[j].line, g_crumbs[j].what);
        break;
      }
  }
  return 0;
}
