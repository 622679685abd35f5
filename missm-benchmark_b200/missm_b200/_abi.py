"""argtypes of every exported entry point of include/missm_b200.h (one table; tests/test_abi.py
checks it against the header and against the built library's symbol table)."""
import ctypes

P = ctypes.c_void_p
I = ctypes.c_int32
L = ctypes.c_int64
F = ctypes.c_float

ABI_VERSION = 9          # == MISSM_ABI_VERSION of include/missm_b200.h; _lib.lib() refuses a library built for another

# name -> argtypes (every function returns int32; 0 = ok)
SIGNATURES = {
    "missm_gemm_bf16": [P, P],
    "missm_gemm_colsum_rows": [I],
    "missm_attention_fwd": [P, P],
    "missm_attention_bwd": [P, P],
    "missm_layernorm_fwd": [P, L, P, P, I, I, P, P, P, P, L, I, P, P, I, I, F, P],
    "missm_ln_bwd_num_partials": [I],
    "missm_layernorm_bwd": [P, L, I, P, L, P, P, P, P, P, P, P, P, P, P, P, I, I, P],
    "missm_reduce_partials": [P, I, L, P, I, F, P],
    "missm_cast_f32_bf16": [P, L, P, L, I, I, I, P],
    "missm_colsum_num_partials": [I],
    "missm_colsum_bf16": [P, L, I, I, P, P, P],
    "missm_patchify": [P, P, P, I, I, I, I, I, I, I, P],
    "missm_patch_embed_implicit": [P, P, P, P, P, I, I, I, I, I, I, I, I, P],
    "missm_colsum_grouped_f32": [P, I, I, I, I, P, P],
    "missm_copy_f32": [P, P, L, P],
    "missm_cls_rows": [P, P, P, I, I, I, P],
    "missm_embed_bwd": [P, P, P, I, I, I, P],
    "missm_frame_mean": [P, P, I, I, I, I, P],
    "missm_frame_mean_bwd": [P, P, I, I, I, P],
    "missm_l2norm_scale_fwd": [P, P, P, F, I, I, P],
    "missm_l2norm_scale_bwd": [P, P, P, F, P, I, I, I, P],
    "missm_text_embed_fwd": [P, P, P, P, P, I, I, I, P],
    "missm_text_embed_bwd": [P, P, P, P, P, I, I, I, P],
    "missm_argmax_rows": [P, P, P, I, I, P],
    "missm_compact_mask": [P, I, P, I, P, P, P, P],
    "missm_scatter_rows_zero": [P, P, P, I, I, P],
    "missm_gather_rows": [P, P, P, I, L, P],
    "missm_fusion_sum_fwd": [P, P],
    "missm_fusion_sum_bwd": [P, P, P, P, P, P],
    "missm_set_persistent_sms": [I],
    "missm_set_coresident": [I],
    "missm_expand6_bf16": [P, L, I, I, P, L, I, I, I, P],
    "missm_patchify_f32": [P, P, P, I, I, I, I, I, I, I, P],
    "missm_gelu_f32_fwd": [P, P, L, P],
    "missm_gelu_f32_bwd": [P, P, P, L, P],
    "missm_attention_f32_fwd": [P, P],
    "missm_attention_f32_bwd": [P, P],
    "missm_adam_multi": [P, P],
    "missm_image_preprocess": [P, P],
    "missm_attn_block_sizes": [P, P],
    "missm_attn_block_fwd": [P, P],
    "missm_attn_block_bwd": [P, P],
    "missm_mlp_block_sizes": [P, P],
    "missm_mlp_block_fwd": [P, P],
    "missm_mlp_block_bwd": [P, P],
    "missm_video_preprocess": [P, P],
    "missm_fbank_num_frames": [L],
    "missm_audio_fbank": [P, P],
    "missm_eval_accumulate": [P, P, I, I, P, P, P, P, P],
    "missm_debug_dump": [],
    "missm_debug_crumb": [ctypes.c_char_p, P],
    "missm_gemm_profile": [I],
    "missm_gemm_profile_read": [P, P, P],
}
# exported but with non-standard return types / no args
OTHER_EXPORTS = ["missm_version", "missm_last_error", "missm_launch_count"]
