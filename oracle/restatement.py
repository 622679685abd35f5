"""TEST INFRASTRUCTURE ONLY -- CPU/fp32 restatement of the reference hot path (the parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product path never does (it must fail loudly without the CUDA library).

What it restates (plain torch, fp32, no autocast, functional over a reference-named state dict):
  * LanguageBind.forward                     languagebind/__init__.py:75-85
  * CLIPVisionTransformer.forward            languagebind/image/modeling_image.py:610-672
                                             (video: languagebind/video/modeling_video.py:723-784)
  * CLIPEncoderLayer.forward                 image/modeling_image.py:86-158 (temporal attn + temporal MLP)
                                             video/modeling_video.py:192-264 (temporal attn only)
  * CLIPTextTransformer.forward              image/modeling_image.py:469-532, causal mask :441-455
  * resize_pos                               image/modeling_image.py:795-839
  * fusion heads + finetune_model            src/model/baseline.py:27-61,65-236,335-453
  * third-party arithmetic (transformers, unpinned, 4.31-4.34 era API; not under /root/reference):
    CLIPAttention, CLIPMLP (quick_gelu), CLIPTextEmbeddings, CLIPVisionEmbeddings (the reference's
    own copy is video/modeling_video.py:19-51), _expand_mask -- restated from their published
    algorithm, anchored on the reference call sites modeling_image.py:121,133,140,150,463,502,602.

Pinning: the reference holds no tests, golden vectors or fixtures, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF executed in the build container
through oracle/ref_shim.py; the vectors live in tests/golden/ and were produced by
oracle/make_golden.py (committed).  tests/test_oracle_golden.py checks this file against them.
"""
import math
import types

import torch
import torch.nn.functional as F

# reference: src/model/baseline.py:8 ; depth / thermal have no code there -- the benchmark
# harness extends the map (5, 6) without touching codes 0-4.
MISSING_TYPE_INDEX = {'language': 1, 'video': 2, 'audio': 3, 'image': 4, 'depth': 5, 'thermal': 6}


def vision_config(**kw):
    """Defaults of CLIPVisionConfig (languagebind/image/configuration_image.py:180-233)."""
    d = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
             num_channels=3, image_size=224, patch_size=32, hidden_act="quick_gelu", layer_norm_eps=1e-5,
             add_time_attn=False, num_frames=1, num_mel_bins=0, target_length=0, temporal_mlp=True,
             lora_r=0, lora_alpha=16)      # lora_r: the reference default is 2 (:200); 0 = plain encoder
    d.update(kw)
    return types.SimpleNamespace(**d)


def text_config(**kw):
    """Defaults of CLIPTextConfig (languagebind/image/configuration_image.py:64-107)."""
    d = dict(vocab_size=49408, hidden_size=512, intermediate_size=2048, num_hidden_layers=12,
             num_attention_heads=8, max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    d.update(kw)
    return types.SimpleNamespace(**d)


def grid_of(cfg):
    """Token grid after resize_pos (modeling_image.py:797-803)."""
    if cfg.num_mel_bins and cfg.target_length:
        size = [int(cfg.num_mel_bins), int(cfg.target_length)]
    else:
        size = cfg.image_size if isinstance(cfg.image_size, (list, tuple)) else [cfg.image_size, cfg.image_size]
    return [size[0] // cfg.patch_size, size[1] // cfg.patch_size]


def resize_pos(old_pos_embed, grid_size):
    """modeling_image.py:795-839: bicubic antialiased resample of the patch position table."""
    new_len = grid_size[0] * grid_size[1] + 1
    if new_len == old_pos_embed.shape[0]:
        return old_pos_embed
    tok, img = old_pos_embed[:1], old_pos_embed[1:]
    og = int(math.sqrt(len(img)))
    img = img.reshape(1, og, og, -1).permute(0, 3, 1, 2)
    img = F.interpolate(img, size=grid_size, mode='bicubic', antialias=True, align_corners=False)
    img = img.permute(0, 2, 3, 1).reshape(1, grid_size[0] * grid_size[1], -1)[0]
    return torch.cat([tok, img], dim=0).to(old_pos_embed.dtype)


def act_fn(name, x):
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu(x)
    raise ValueError(name)


def lora_linear(sd, pre, x, scaling):
    """A Linear that peft may have wrapped (convert_to_lora, modeling_image.py:775-793; peft is third-party,
    unpinned and absent -- restated from its published algorithm): W x + b + (lora_alpha / r) B(A(x)).
    lora_dropout is the identity at the reference default 0.0 (configuration_image.py:202) and in eval mode."""
    y = F.linear(x, sd[pre + 'weight'], sd.get(pre + 'bias'))
    a = sd.get(pre + 'lora_A.default.weight')
    if a is not None:
        y = y + F.linear(F.linear(x, a), sd[pre + 'lora_B.default.weight']) * scaling
    return y


def clip_attention(sd, pre, x, n_heads, causal_mask=None, attn_mask=None, lora_scaling=1.0):
    """transformers 4.3x CLIPAttention.forward."""
    b, n, d = x.shape
    hd = d // n_heads
    q = lora_linear(sd, pre + 'q_proj.', x, lora_scaling) * hd ** -0.5
    k = lora_linear(sd, pre + 'k_proj.', x, lora_scaling)
    v = lora_linear(sd, pre + 'v_proj.', x, lora_scaling)
    sh = lambda t: t.view(b, n, n_heads, hd).transpose(1, 2)
    w = sh(q) @ sh(k).transpose(-1, -2)                      # [b, H, n, n]
    if causal_mask is not None:
        w = w + causal_mask
    if attn_mask is not None:
        w = w + attn_mask
    w = torch.softmax(w, dim=-1)
    o = (w @ sh(v)).transpose(1, 2).reshape(b, n, d)
    return lora_linear(sd, pre + 'out_proj.', o, lora_scaling)


def clip_mlp(sd, pre, x, act, lora_scaling=1.0):
    """transformers CLIPMLP.forward: fc2(act(fc1(x))) (temporal_mlp.fc1 / fc2 are LoRA targets, modeling_image.py:780-781)."""
    h = act_fn(act, lora_linear(sd, pre + 'fc1.', x, lora_scaling))
    return lora_linear(sd, pre + 'fc2.', h, lora_scaling)


def layer_norm(sd, pre, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + 'weight'], sd[pre + 'bias'], eps)


def encoder_layer(sd, pre, x, cfg, causal_mask=None, attn_mask=None):
    """CLIPEncoderLayer.forward (modeling_image.py:86-158; video :192-264)."""
    H, eps, act = cfg.num_attention_heads, cfg.layer_norm_eps, cfg.hidden_act
    ls = cfg.lora_alpha / cfg.lora_r if getattr(cfg, 'lora_r', 0) else 1.0
    if getattr(cfg, 'add_time_attn', False):
        bt, n, d = x.shape
        t = cfg.num_frames
        to_time = lambda z: z.view(bt // t, t, n, d).permute(0, 2, 1, 3).reshape(-1, t, d)   # (b t) n d -> (b n) t d
        to_space = lambda z: z.view(bt // t, n, t, d).permute(0, 2, 1, 3).reshape(bt, n, d)  # (b n) t d -> (b t) n d
        if t != 1:
            x = to_space(to_time(x) + sd[pre + 'temporal_embedding'][:, :t, :])
        res = x
        h = layer_norm(sd, pre + 'temporal_layer_norm1.', to_time(x), eps)
        h = clip_attention(sd, pre + 'temporal_attn.', h, H, causal_mask, attn_mask, ls)
        x = res + to_space(h)
        if cfg.temporal_mlp:          # image/audio/depth/thermal variant; commented out for video
            res = x
            h = layer_norm(sd, pre + 'temporal_layer_norm2.', to_time(x), eps)
            h = clip_mlp(sd, pre + 'temporal_mlp.', h, act, ls)
            x = res + to_space(h)
    res = x
    h = layer_norm(sd, pre + 'layer_norm1.', x, eps)
    x = res + clip_attention(sd, pre + 'self_attn.', h, H, causal_mask, attn_mask, ls)
    res = x
    h = layer_norm(sd, pre + 'layer_norm2.', x, eps)
    return res + clip_mlp(sd, pre + 'mlp.', h, act)


def vision_tower(sd, pre, pixel_values, cfg):
    """CLIPVisionTransformer.forward (modeling_image.py:610-672) -> pooled [B, D]."""
    if pixel_values.dim() == 5:                                      # b c t h w -> (b t) c h w
        B, C, T, Hh, Ww = pixel_values.shape
        pixel_values = pixel_values.permute(0, 2, 1, 3, 4).reshape(B * T, C, Hh, Ww)
    else:
        B, T = pixel_values.shape[0], 1
    # CLIPVisionEmbeddings (video/modeling_video.py:42-51)
    pe = F.conv2d(pixel_values, sd[pre + 'embeddings.patch_embedding.weight'], stride=cfg.patch_size)
    pe = pe.flatten(2).transpose(1, 2)
    cls = sd[pre + 'embeddings.class_embedding'].expand(pe.shape[0], 1, -1)
    x = torch.cat([cls, pe], dim=1) + sd[pre + 'embeddings.position_embedding.weight'][None]
    # PatchDropout is the identity at force_patch_dropout = 0 (modeling_image.py:31-32)
    x = layer_norm(sd, pre + 'pre_layrnorm.', x, cfg.layer_norm_eps)
    # a peft-wrapped encoder (lora_r != 0) keeps its layers under encoder.base_model.model (modeling_image.py:793)
    enc = f'{pre}encoder.base_model.model.' if getattr(cfg, 'lora_r', 0) else f'{pre}encoder.'
    for i in range(cfg.num_hidden_layers):
        x = encoder_layer(sd, f'{enc}layers.{i}.', x, cfg)
    pooled = layer_norm(sd, pre + 'post_layernorm.', x[:, 0, :], cfg.layer_norm_eps)
    return pooled.reshape(B, T, -1).mean(1)


def text_tower(sd, pre, input_ids, attention_mask, cfg):
    """CLIPTextTransformer.forward (modeling_image.py:469-532) -> pooled [B, D]."""
    B, L = input_ids.shape
    x = sd[pre + 'embeddings.token_embedding.weight'][input_ids] + \
        sd[pre + 'embeddings.position_embedding.weight'][:L][None]
    fmin = torch.finfo(x.dtype).min
    causal = torch.full((L, L), fmin).triu(1)[None, None]            # _make_causal_mask :441-455
    amask = None
    if attention_mask is not None:                                   # _expand_mask
        inv = 1.0 - attention_mask[:, None, None, :].expand(B, 1, L, L).to(x.dtype)
        amask = inv.masked_fill(inv.to(torch.bool), fmin)
    for i in range(cfg.num_hidden_layers):
        x = encoder_layer(sd, f'{pre}encoder.layers.{i}.', x, cfg, causal, amask)
    x = layer_norm(sd, pre + 'final_layer_norm.', x, cfg.layer_norm_eps)
    return x[torch.arange(B), input_ids.to(torch.int).argmax(dim=-1)]   # :519-522


def languagebind_forward(sd, inputs, cfgs, text_cfg, logit_scales, use_temp=True, pre='modality_'):
    """LanguageBind.forward (languagebind/__init__.py:75-85)."""
    out = {}
    for key, value in inputs.items():
        if key == 'language':
            v = text_tower(sd, f'{pre}encoder.language.', value['input_ids'], value.get('attention_mask'), text_cfg)
        else:
            v = vision_tower(sd, f'{pre}encoder.{key}.', value['pixel_values'], cfgs[key])
        v = F.linear(v, sd[f'{pre}proj.{key}.weight'])
        v = v / v.norm(p=2, dim=-1, keepdim=True)
        if use_temp and key != 'language':
            v = v * math.exp(float(logit_scales[key]))
        out[key] = v
    return out


# ------------------------------------------------------------------------------ fusion heads
def _head(sd, pre, x):
    """Head (baseline.py:27-39) in eval mode (dropout = identity)."""
    h = F.relu(F.linear(x, sd[pre + 'head.0.weight'], sd[pre + 'head.0.bias']))
    return F.linear(h, sd[pre + 'head.3.weight'], sd[pre + 'head.3.bias'])


def _norm(sd, pre, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + 'norm.weight'], sd[pre + 'norm.bias'], 1e-5)


def _proj(sd, pre, modal, x):
    return F.linear(x, sd[f'{pre}modal_proj.{modal}.weight'], sd[f'{pre}modal_proj.{modal}.bias'])


def fusion_forward(sd, fusion_type, modality_types, batch, missing_index, pre='fusion.', training=False):
    """Eval-mode fusion heads of src/model/baseline.py (dropout inactive)."""
    miss = {m: missing_index == MISSING_TYPE_INDEX[m] for m in modality_types}
    batch = {k: v.clone() for k, v in batch.items()}        # the reference mutates in place
    if fusion_type == 'sum':                                # :52-61
        acc = 0
        for m in modality_types:
            d = _proj(sd, pre, m, batch[m])
            d[miss[m]] = 0
            acc = acc + d
        return _head(sd, pre + 'head.', _norm(sd, pre, acc))
    if fusion_type in ('concat', 'retrieval'):              # :77-86, :162-169
        ins = []
        for m in modality_types:
            if fusion_type == 'concat' and miss[m].any():
                batch[m][miss[m]] = sd[f'{pre}statistics_{m}']
            ins.append(_proj(sd, pre, m, batch[m]))
        return _head(sd, pre + 'head.', _norm(sd, pre, torch.cat(ins, -1)))
    if fusion_type == 'regression':                         # :112-149
        pf = {m: _proj(sd, pre, m, batch[m]) for m in modality_types}
        for tm in modality_types:
            if not miss[tm].any():
                continue
            preds, masks = [], []
            for sm in modality_types:
                if sm != tm:
                    k = f'{pre}cross_modal_regressors.{sm}_to_{tm}.'
                    preds.append(F.linear(batch[sm], sd[k + 'weight'], sd[k + 'bias']))
                    masks.append((~miss[sm]).float())
            preds = torch.stack(preds, 1)
            masks = torch.stack(masks, -1).unsqueeze(-1)
            avg = (preds * masks).sum(1) / masks.sum(1).clamp(min=1e-6)
            filled = pf[tm].clone()
            filled[miss[tm]] = avg[miss[tm]]
            pf[tm] = filled
        return _head(sd, pre + 'head.', _norm(sd, pre, torch.cat([pf[m] for m in modality_types], -1)))
    if fusion_type == 'intra_attention':                    # :191-203
        acc = 0
        for m in modality_types:
            d = _proj(sd, pre, m, batch[m])
            rep = sd[pre + 'fusion_representation'].expand(d.shape[0], -1)
            ca = F.linear(torch.cat([d, rep], -1), sd[pre + 'channel_attention.0.weight'], sd[pre + 'channel_attention.0.bias'])
            ca = torch.sigmoid(F.linear(F.relu(ca), sd[pre + 'channel_attention.2.weight'], sd[pre + 'channel_attention.2.bias']))
            d = d * ca
            d[miss[m]] = 0
            acc = acc + d
        return _head(sd, pre + 'head.', _norm(sd, pre, acc))
    if fusion_type == 'inter_attention':                    # :220-236  (nn.MultiheadAttention, 4 heads)
        toks = torch.stack([_proj(sd, pre, m, batch[m]) for m in modality_types], 1)   # [B, M, D]
        mask = torch.stack([miss[m] for m in modality_types], 1)
        B, M, D = toks.shape
        H, hd = 4, D // 4
        W, bI = sd[pre + 'attn.in_proj_weight'], sd[pre + 'attn.in_proj_bias']
        q = F.linear(sd[pre + 'query_token'].expand(B, -1, -1), W[:D], bI[:D])
        k = F.linear(toks, W[D:2 * D], bI[D:2 * D])
        v = F.linear(toks, W[2 * D:], bI[2 * D:])
        sh = lambda t: t.view(B, -1, H, hd).transpose(1, 2)
        s = (sh(q) * hd ** -0.5) @ sh(k).transpose(-1, -2)
        s = s.masked_fill(mask[:, None, None, :], float('-inf'))
        o = (torch.softmax(s, -1) @ sh(v)).transpose(1, 2).reshape(B, 1, D)
        o = F.linear(o, sd[pre + 'attn.out_proj.weight'], sd[pre + 'attn.out_proj.bias'])
        return _head(sd, pre + 'head.', _norm(sd, pre, o[:, 0]))
    if fusion_type == 'dedicated_dnn':                      # :346-354
        feats = torch.stack([batch[m] for m in modality_types], 1)
        B = feats.shape[0]
        out = F.linear(feats.view(B, -1), sd[pre + 'dedicated_dnn.full.weight'], sd[pre + 'dedicated_dnn.full.bias'])
        for i, m in enumerate(modality_types):
            alt = F.linear(torch.cat([feats[:, :i], feats[:, i + 1:]], 1).view(B, -1),
                           sd[f'{pre}dedicated_dnn.{m}.weight'], sd[f'{pre}dedicated_dnn.{m}.bias'])
            out[miss[m]] = alt[miss[m]]
        return _head(sd, pre + 'head.', _norm(sd, pre, out))
    if fusion_type in ('Distill_tea', 'MTD_stu', 'KL_stu', 'self_distill'):   # :371-380, :397-418
        mp = lambda z: F.linear(F.relu(F.linear(z, sd[pre + 'modal_proj.0.weight'], sd[pre + 'modal_proj.0.bias'])),
                                sd[pre + 'modal_proj.2.weight'], sd[pre + 'modal_proj.2.bias'])
        for m in modality_types:
            batch[m][miss[m]] = 0
        feats = torch.cat([batch[m] for m in modality_types], -1)
        if fusion_type == 'self_distill':
            if training:
                B, C = batch[modality_types[0]].shape
                stu = []
                for i, m in enumerate(modality_types):
                    z = torch.cat([torch.zeros(B, i * C), batch[m], torch.zeros(B, (len(modality_types) - i - 1) * C)], -1)
                    stu.append(mp(z))
                tea = mp(feats)
                return [~miss[m] for m in modality_types], stu, tea, _head(sd, pre + 'head.', _norm(sd, pre, tea))
            return _head(sd, pre + 'head.', _norm(sd, pre, mp(feats)))
        return feats, _head(sd, pre + 'head.', _norm(sd, pre, mp(feats)))
    if fusion_type == 'graph_fusion':                       # :254-268
        x = torch.stack([_proj(sd, pre, m, batch[m]) for m in modality_types], 1)
        outs = []
        for b in range(x.shape[0]):                         # one graph per sample (Batch.from_data_list keeps them apart)
            present = [not bool(miss[m][b]) for m in modality_types]
            outs.append(fusion_gcn(sd, pre + 'gcn.', x[b], build_edge(present)).mean(0))
        return _head(sd, pre + 'head.', _norm(sd, pre, torch.stack(outs)))
    if fusion_type == 'unified_graph':                      # :298-329
        feats = torch.stack([batch[m] for m in modality_types], 1)
        outs = []
        for b in range(feats.shape[0]):
            present = [not bool(miss[m][b]) for m in modality_types]
            done = fusion_gcn(sd, pre + 'complete_gcn.', feats[b], build_edge(present), hidden=384, heads=4)
            f = torch.stack([feats[b, i] if present[i] else done[i] for i in range(len(modality_types))])
            outs.append(fusion_gcn(sd, pre + 'fusion_gcn.', f, build_edge([True] * len(modality_types))).mean(0))
        return _head(sd, pre + 'head.', _norm(sd, pre, torch.stack(outs)))
    raise ValueError(f"unknown fusion type {fusion_type!r}")


def build_edge(present):
    """bulid_edge (baseline.py:270-281): both directions of every pair of PRESENT nodes -> int64 [2, E]."""
    start, end = [], []
    for i in range(len(present)):
        for j in range(i + 1, len(present)):
            if present[i] and present[j]:
                start.append(i)
                end.append(j)
    return torch.tensor([start + end, end + start], dtype=torch.long)


def super_gat_conv(sd, pre, x, edge_index, heads, concat, negative_slope=0.2):
    """torch_geometric.nn.SuperGATConv.forward (third party, unpinned, not installed; call sites baseline.py:14-15,
    20,22), attention_type='MX', add_self_loops=True, dropout 0 -- restated from its published algorithm as the
    edge-list / scatter-softmax computation PyG performs.  The self-supervised attention loss it also prepares in
    training mode (att_x / att_y) is never read by the reference."""
    N = x.shape[0]
    w = sd[pre + 'lin.weight']
    C = w.shape[0] // heads
    h = F.linear(x, w).view(N, heads, C)
    keep = edge_index[0] != edge_index[1]                   # remove_self_loops, then add_self_loops
    loops = torch.arange(N)
    src = torch.cat([edge_index[0][keep], loops])
    dst = torch.cat([edge_index[1][keep], loops])
    x_j, x_i = h[src], h[dst]
    logits = (x_i * x_j).sum(-1)
    alpha = (x_j * sd[pre + 'att_l']).sum(-1) + (x_i * sd[pre + 'att_r']).sum(-1)
    alpha = F.leaky_relu(alpha * logits.sigmoid(), negative_slope)
    out = torch.zeros_like(h)
    for i in range(N):                                      # softmax over the incoming edges of node i, then aggregate
        e = dst == i
        a = torch.softmax(alpha[e], dim=0)
        out[i] = (a.unsqueeze(-1) * x_j[e]).sum(0)
    out = out.reshape(N, heads * C) if concat else out.mean(1)
    return out + sd[pre + 'bias']


def fusion_gcn(sd, pre, x, edge_index, hidden=128, heads=4):
    """fusion_gcn.forward (baseline.py:11-24): SuperGATConv(heads=4, concat) -> GELU -> SuperGATConv(heads=1, mean)."""
    x = super_gat_conv(sd, pre + 'gat1.', x, edge_index, heads, True)
    return super_gat_conv(sd, pre + 'gat2.', F.gelu(x), edge_index, 1, False)


def finetune_forward(sd, fusion_type, modality_types, data, missing_index, cfgs, text_cfg, logit_scales,
                     use_temp=True):
    """finetune_model.forward (src/model/baseline.py:450-453) on a full state dict with the
    reference's key names (`encoder.modality_encoder...`, `fusion...`)."""
    emb = languagebind_forward(sd, data, cfgs, text_cfg, logit_scales, use_temp, pre='encoder.modality_')
    return fusion_forward(sd, fusion_type, modality_types, emb, missing_index), emb


# ------------------------------------------------------------------ input pipeline (image-shaped modalities)
OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)      # languagebind/image/processing_image.py:10-11
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)


def image_transform(img_u8_hwc, size=224, antialias=True):
    """get_image_transform / get_thermal_transform (processing_image.py:20-29, thermal/processing_thermal.py:15-25):
    ToTensor -> Resize(224, BICUBIC) -> CenterCrop(224) -> Normalize, on a decoded RGB image (numpy uint8 [H, W, 3]).
    torchvision is the third party the reference calls (unpinned): `antialias` is its Resize default on tensors,
    False up to 0.16 (the reference's era), True from 0.17 (what the reference computes in this image)."""
    import numpy as np
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF
    t = torch.from_numpy(np.ascontiguousarray(img_u8_hwc)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    t = TF.resize(t, size, interpolation=InterpolationMode.BICUBIC, antialias=antialias)
    t = TF.center_crop(t, size)
    return TF.normalize(t, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD)


def depth_transform(depth_f32_hw, max_depth=10.0, size=224, antialias=True):
    """get_depth_transform (depth/processing_depth.py:21-57): DepthNorm (/1000, clip [0.01, max_depth], / max_depth,
    one channel repeated to three), then the image chain."""
    import numpy as np
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF
    d = depth_f32_hw.astype(np.float32) / 1000.0
    d = d.clip(min=0.01)
    d = d.clip(max=max_depth)
    d /= max_depth
    t = torch.from_numpy(d).unsqueeze(0).repeat(3, 1, 1).to(torch.float32)
    t = TF.resize(t, size, interpolation=InterpolationMode.BICUBIC, antialias=antialias)
    t = TF.center_crop(t, size)
    return TF.normalize(t, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD)


def synth_image(seed, H, W):
    """Deterministic decoded 'photo': smooth gradients + noise, uint8 [H, W, 3] (numpy Generator, platform independent)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    base = np.stack([127 + 100 * np.sin(xx / 17.0 + c) * np.cos(yy / 23.0 - c) for c in range(3)], -1)
    return np.clip(base + rng.integers(-40, 41, (H, W, 3)), 0, 255).astype(np.uint8)


def synth_depth(seed, H, W):
    """Deterministic 16-bit depth map in millimetres (0 = holes, up to 12 m so that both clips act), float32 [H, W]."""
    import numpy as np
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 12000, (H, W)).astype(np.float32)
    d[rng.random((H, W)) < 0.05] = 0.0
    return d


# ------------------------------------------------------------------ deterministic synthetic setup
def synth_param(name, shape, std):
    """Order-independent synthetic weights: each tensor is drawn from its own generator seeded by
    a stable hash of its NAME, so the reference module tree, this oracle and the CUDA modules
    all get the identical values regardless of construction order."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def synth_state_dict(named_shapes):
    """named_shapes: iterable of (name, shape).  LayerNorm weights ~ 1 + 0.1 n, biases 0.02 n,
    matrices ~ N(0, fan_in^-1/2 * 0.8) so activations stay O(1) through 24 layers."""
    sd = {}
    for name, shape in named_shapes:
        shape = tuple(shape)
        leaf = name.rsplit('.', 1)[-1]
        if 'position_ids' in name:
            continue
        if ('norm' in name or 'layrnorm' in name) and leaf == 'weight':
            sd[name] = 1.0 + synth_param(name, shape, 0.1)
        elif leaf == 'bias' or 'in_proj_bias' in name:
            sd[name] = synth_param(name, shape, 0.02)
        elif 'lora_B' in name:
            # peft initialises B = 0 (the adapter is a no-op); a trained adapter is what needs checking
            sd[name] = synth_param(name, shape, 0.05 * shape[1] ** -0.5)
        elif 'statistics_' in name:
            sd[name] = synth_param(name, shape, 0.05)
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if 'embedding' in name and 'patch' not in name:
                sd[name] = synth_param(name, shape, 0.02 if 'token' in name or 'position' in name else 0.03)
            else:
                sd[name] = synth_param(name, shape, 0.8 * fan_in ** -0.5)
        elif len(shape) == 0:
            sd[name] = torch.tensor(2.6592)
        else:
            sd[name] = synth_param(name, shape, 0.03)
    return sd


def synth_inputs(modalities, B, cfgs, text_cfg, seed=0):
    """Synthetic inputs in the loader's contract."""
    g = torch.Generator().manual_seed(seed)
    data = {}
    for m in modalities:
        if m == 'language':
            L = text_cfg.max_position_embeddings
            V = text_cfg.vocab_size
            ids = torch.full((B, L), V - 1, dtype=torch.long)          # EOT padding (= max id)
            ids[:, 0] = V - 2                                            # BOS
            n_body = min(19, L - 2)
            ids[:, 1:1 + n_body] = torch.randint(1, max(2, V - 408), (B, n_body), generator=g)
            am = torch.zeros((B, L), dtype=torch.long)
            am[:, :n_body + 2] = 1
            data[m] = {'input_ids': ids, 'attention_mask': am}
        else:
            c = cfgs[m]
            gh, gw = grid_of(c)
            Hh, Ww = gh * c.patch_size, gw * c.patch_size
            if getattr(c, 'add_time_attn', False) and c.num_frames > 1:
                px = torch.randn((B, 3, c.num_frames, Hh, Ww), generator=g)
            else:
                px = torch.randn((B, 3, Hh, Ww), generator=g)
            data[m] = {'pixel_values': px}
    return data


def synth_missing_index(B, ratio, modality_types, seed=2025):
    """src/utils/generate_missing.py:23-38 ("mixed"): int(B*ratio) samples by random.sample, each
    assigned one code uniformly from the towers in use."""
    import random
    rng = random.Random(seed)
    mi = [0] * B
    for i in rng.sample(range(B), int(B * ratio)):
        mi[i] = MISSING_TYPE_INDEX[rng.choice(list(modality_types))]
    return torch.tensor(mi, dtype=torch.long)
