"""Drop-in `src.model.baseline`: `finetune_model` and the missing-modality fusion heads with the
reference's parameter names (so `final_model/<ds>_<fusion>.pth` files load, test.py:92), the
same constructor arguments (`args.fusion_type, modality_types, feature_dims, fusion_dim,
dropout_prob`), the same return arity per fusion type (train_ddp.py:109-112,232-249) and the
`fusion.set_statistics` hook (test.py:115).

Reference: src/model/baseline.py -- Head :27-39, modal_sum :43-61, modal_concat :65-90,
modal_regression :94-149, modal_concat_full :153-169, modal_intra_channel_attention :173-203,
modal_inter_attention :207-236, modal_dedicated_dnn :335-354, modal_distillation :358-380,
modal_self_distillation :384-418, finetune_model :421-453; fusion_gcn :11-24, modal_graph_fusion :240-281,
modal_unified_graph :285-331 (SuperGATConv of torch_geometric -- third party, unpinned, not installed -- is restated
below from its published algorithm as a dense per-sample computation over the <= 6 modality nodes; these two heads are
NOT skip-safe -- a missing modality's embedding still enters through its self-loop -- so the towers run the full batch
for them).

The towers (>99.9 % of the step) run the hand-written CUDA path; `finetune_model.forward` hands
`missing_index` to the encoder bank so that a tower only computes its PRESENT samples.  The
default head (`sum`) runs the fused masked-fusion CUDA kernels of csrc/fusion.cu; the other heads
are a handful of [B, <=1536] tensor ops expressed with torch.nn (library kernels).
"""
import torch
from torch import nn

from languagebind import LanguageBind, to_device, transform_dict, LanguageBindImageTokenizer  # noqa: F401
from missm_b200.bank import MISSING_TYPE_INDEX
from missm_b200 import fusion_ops

missing_type_index = MISSING_TYPE_INDEX


def _miss(missing_index, modal):
    return missing_index == missing_type_index[modal]


class Head(nn.Module):
    def __init__(self, args, input_dims, output_dims):
        super().__init__()
        self.head = nn.Sequential(
            nn.Linear(input_dims, args.fusion_dim),
            nn.ReLU(inplace=True),
            nn.Dropout(args.dropout_prob),
            nn.Linear(args.fusion_dim, output_dims),
        )

    def forward(self, inputs):
        return self.head(inputs)


class modal_sum(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        self.modal_proj = nn.ModuleDict({m: nn.Linear(args.feature_dims, args.fusion_dim) for m in args.modality_types})
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)

    def forward(self, batch, missing_index):
        # fused: per-modality Linear, zero rows of missing samples, sum, LayerNorm  (:52-61)
        fused = fusion_ops.masked_sum_norm(
            [batch[m] for m in self.modality_types],
            [self.modal_proj[m].weight for m in self.modality_types],
            [self.modal_proj[m].bias for m in self.modality_types],
            [missing_type_index[m] for m in self.modality_types],
            missing_index, self.norm.weight, self.norm.bias, self.norm.eps)
        return self.head(fused)


class modal_concat(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        n = len(args.modality_types)
        self.modal_proj = nn.ModuleDict({m: nn.Linear(args.feature_dims, args.fusion_dim) for m in args.modality_types})
        self.norm = nn.LayerNorm(args.fusion_dim * n)
        self.head = Head(args, args.fusion_dim * n, output_dims)
        for m in self.modality_types:
            self.register_buffer(f'statistics_{m}', torch.zeros(args.feature_dims, dtype=torch.float))

    def forward(self, batch, missing_index):
        inputs = []
        for m in self.modality_types:
            stat = self.get_buffer(f'statistics_{m}').to(batch[m].device)
            x = torch.where(_miss(missing_index, m)[:, None], stat[None, :].to(batch[m].dtype), batch[m])
            inputs.append(self.modal_proj[m](x))
        return self.head(self.norm(torch.cat(inputs, dim=-1)))

    def set_statistics(self, statistics, modality_types):
        for m in modality_types:
            self.register_buffer(f'statistics_{m}', torch.as_tensor(statistics[m], dtype=torch.float,
                                                                    device=self.modal_proj[m].weight.device))


class modal_regression(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        n = len(args.modality_types)
        self.modal_proj = nn.ModuleDict({m: nn.Linear(args.feature_dims, args.fusion_dim) for m in args.modality_types})
        self.norm = nn.LayerNorm(args.fusion_dim * n)
        self.head = Head(args, args.fusion_dim * n, output_dims)
        self.cross_modal_regressors = nn.ModuleDict()
        for s in self.modality_types:
            for t in self.modality_types:
                if s != t:
                    self.cross_modal_regressors[f"{s}_to_{t}"] = nn.Linear(args.feature_dims, args.fusion_dim)

    def forward(self, batch, missing_index):
        pf = {m: self.modal_proj[m](batch[m]) for m in self.modality_types}
        for t in self.modality_types:
            tmask = _miss(missing_index, t)
            preds, masks = [], []
            for s in self.modality_types:
                if s != t:
                    preds.append(self.cross_modal_regressors[f"{s}_to_{t}"](batch[s]))
                    masks.append((~_miss(missing_index, s)).to(preds[-1].dtype))
            preds = torch.stack(preds, dim=1)
            masks = torch.stack(masks, dim=-1).unsqueeze(-1)
            avg = (preds * masks).sum(dim=1) / masks.sum(dim=1).clamp(min=1e-6)
            pf[t] = torch.where(tmask[:, None], avg, pf[t])
        return self.head(self.norm(torch.cat([pf[m] for m in self.modality_types], dim=-1)))


class modal_concat_full(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        n = len(args.modality_types)
        self.modal_proj = nn.ModuleDict({m: nn.Linear(args.feature_dims, args.fusion_dim) for m in args.modality_types})
        self.norm = nn.LayerNorm(args.fusion_dim * n)
        self.head = Head(args, args.fusion_dim * n, output_dims)

    def forward(self, batch, missing_index):
        return self.head(self.norm(torch.cat([self.modal_proj[m](batch[m]) for m in self.modality_types], dim=-1)))


class modal_intra_channel_attention(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        self.modal_proj = nn.ModuleDict({m: nn.Linear(args.feature_dims, args.fusion_dim) for m in args.modality_types})
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)
        self.fusion_representation = nn.Parameter(torch.randn(1, args.fusion_dim))
        self.channel_attention = nn.Sequential(
            nn.Linear(args.fusion_dim * 2, args.fusion_dim // 4), nn.ReLU(),
            nn.Linear(args.fusion_dim // 4, args.fusion_dim), nn.Sigmoid())

    def forward(self, batch, missing_index):
        total = 0
        for m in self.modality_types:
            d = self.modal_proj[m](batch[m])
            ca = self.channel_attention(torch.cat([d, self.fusion_representation.expand(d.shape[0], -1)], dim=-1))
            d = d * ca
            total = total + torch.where(_miss(missing_index, m)[:, None], torch.zeros_like(d), d)
        return self.head(self.norm(total))


class modal_inter_attention(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        self.modal_proj = nn.ModuleDict({m: nn.Linear(args.feature_dims, args.fusion_dim) for m in args.modality_types})
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)
        self.query_token = nn.Parameter(torch.randn(1, 1, args.fusion_dim))
        self.attn = nn.MultiheadAttention(args.fusion_dim, num_heads=4, batch_first=True)

    def forward(self, batch, missing_index):
        toks = torch.stack([self.modal_proj[m](batch[m]) for m in self.modality_types], dim=1)
        mask = torch.stack([_miss(missing_index, m) for m in self.modality_types], dim=1)
        query = self.query_token.expand(toks.shape[0], -1, -1)
        out, _ = self.attn(query, toks, toks, key_padding_mask=mask.bool())
        return self.head(self.norm(out[:, 0, :]))


class modal_dedicated_dnn(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        n = len(self.modality_types)
        d = {m: nn.Linear(args.feature_dims * (n - 1), args.fusion_dim) for m in args.modality_types}
        d['full'] = nn.Linear(args.feature_dims * n, args.fusion_dim)
        self.dedicated_dnn = nn.ModuleDict(d)
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)

    def forward(self, batch, missing_index):
        feats = torch.stack([batch[m] for m in self.modality_types], dim=1)
        B = feats.shape[0]
        out = self.dedicated_dnn['full'](feats.view(B, -1))
        for i, m in enumerate(self.modality_types):
            alt = self.dedicated_dnn[m](torch.cat([feats[:, :i], feats[:, i + 1:]], dim=1).view(B, -1))
            out = torch.where(_miss(missing_index, m)[:, None], alt, out)
        return self.head(self.norm(out))


def _zeroed(batch, missing_index, modality_types):
    return [torch.where(_miss(missing_index, m)[:, None], torch.zeros_like(batch[m]), batch[m])
            for m in modality_types]


class modal_distillation(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        n = len(self.modality_types)
        self.modal_proj = nn.Sequential(nn.Linear(args.feature_dims * n, args.fusion_dim), nn.ReLU(),
                                        nn.Linear(args.fusion_dim, args.fusion_dim))
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)

    def forward(self, batch, missing_index):
        features = torch.cat(_zeroed(batch, missing_index, self.modality_types), dim=-1)
        return features, self.head(self.norm(self.modal_proj(features)))


class modal_self_distillation(nn.Module):
    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        n = len(self.modality_types)
        self.modal_proj = nn.Sequential(nn.Linear(args.feature_dims * n, args.fusion_dim), nn.ReLU(),
                                        nn.Linear(args.fusion_dim, args.fusion_dim))
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)

    def forward(self, batch, missing_index):
        ori = _zeroed(batch, missing_index, self.modality_types)
        if not self.training:
            return self.head(self.norm(self.modal_proj(torch.cat(ori, dim=-1))))
        B, Cd = ori[0].shape
        n = len(self.modality_types)
        stu, masks = [], []
        for i, m in enumerate(self.modality_types):
            z = torch.cat([ori[i].new_zeros((B, i * Cd)), ori[i], ori[i].new_zeros((B, (n - i - 1) * Cd))], dim=-1)
            stu.append(self.modal_proj(z))
            masks.append(missing_index != missing_type_index[m])
        tea = self.modal_proj(torch.cat(ori, dim=-1))
        return masks, stu, tea, self.head(self.norm(tea))


# ------------------------------------------------------------------------------------------------
# Graph heads (reference :11-24, :240-331).  torch_geometric.nn.SuperGATConv (attention_type 'MX', the default):
#   h = lin(x) viewed [N, H, C];  for an edge j -> i (self-loops added for every node):
#   e_ij = leaky_relu( ((h_j . att_l) + (h_i . att_r)) * sigmoid(h_i . h_j), 0.2 );  alpha_i. = softmax_j(e_ij)
#   out_i = sum_j alpha_ij h_j;  heads concatenated (concat=True) or averaged;  + bias.
# Its self-supervised edge loss (att_x / att_y, negative sampling) is never read by the reference and is not built.
# The graphs here are one per sample over the M = len(modality_types) nodes, so the edge lists become a dense
# [B, M, M] adjacency mask and the scatter-softmax a masked softmax.
# ------------------------------------------------------------------------------------------------
class _PygLinear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear(bias=False): parameter name `weight` [out, in], glorot init."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight)


class SuperGATConv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2):
        super().__init__()
        self.heads, self.out_channels, self.concat, self.negative_slope = heads, out_channels, concat, negative_slope
        self.att_l = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_r = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        self.lin = _PygLinear(in_channels, heads * out_channels)
        nn.init.xavier_uniform_(self.att_l)
        nn.init.xavier_uniform_(self.att_r)

    def forward(self, x, adj):
        """x [B, M, in]; adj bool [B, M, M], adj[b, i, j] = edge j -> i of sample b (self-loops included)."""
        B, M, _ = x.shape
        h = self.lin(x).view(B, M, self.heads, self.out_channels)
        logits = torch.einsum('bihc,bjhc->bhij', h, h)
        src = (h * self.att_l).sum(-1).permute(0, 2, 1)                  # [B, H, j]
        dst = (h * self.att_r).sum(-1).permute(0, 2, 1)                  # [B, H, i]
        e = (src[:, :, None, :] + dst[:, :, :, None]) * logits.sigmoid()
        e = torch.nn.functional.leaky_relu(e, self.negative_slope)
        alpha = e.masked_fill(~adj[:, None], float('-inf')).softmax(dim=-1)
        out = torch.einsum('bhij,bjhc->bihc', alpha, h)
        out = out.reshape(B, M, self.heads * self.out_channels) if self.concat else out.mean(dim=2)
        return out + self.bias


class fusion_gcn(nn.Module):
    def __init__(self, in_channels=256, hidden_dim=128, output_dim=256, heads=4):
        super().__init__()
        self.gat1 = SuperGATConv(in_channels, hidden_dim, heads=heads, concat=True)
        self.gat2 = SuperGATConv(hidden_dim * heads, output_dim, heads=1, concat=False)
        self.act = nn.GELU()

    def forward(self, x, adj):
        return self.gat2(self.act(self.gat1(x, adj)), adj)


def _adjacency(present):
    """bulid_edge (:270-281) + SuperGATConv's self-loops: i <-> j when BOTH modalities are present, i -> i always."""
    M = present.shape[1]
    both = present[:, :, None] & present[:, None, :]
    return both | torch.eye(M, dtype=torch.bool, device=present.device)[None]


class modal_graph_fusion(nn.Module):
    compaction_safe = False

    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        self.modal_proj = nn.ModuleDict({m: nn.Linear(args.feature_dims, args.fusion_dim) for m in args.modality_types})
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)
        self.gcn = fusion_gcn()

    def forward(self, batch, missing_index):
        x = torch.stack([self.modal_proj[m](batch[m]) for m in self.modality_types], dim=1)
        present = torch.stack([~_miss(missing_index, m) for m in self.modality_types], dim=1)
        out = self.gcn(x, _adjacency(present)).mean(dim=-2)
        return self.head(self.norm(out))


class modal_unified_graph(nn.Module):
    compaction_safe = False

    def __init__(self, args, output_dims):
        super().__init__()
        self.modality_types = args.modality_types
        self.norm = nn.LayerNorm(args.fusion_dim)
        self.head = Head(args, args.fusion_dim, output_dims)
        self.complete_gcn = fusion_gcn(in_channels=768, hidden_dim=384, output_dim=768)
        self.fusion_gcn = fusion_gcn(in_channels=768)

    def forward(self, batch, missing_index):
        feats = torch.stack([batch[m] for m in self.modality_types], dim=1)
        missing = torch.stack([_miss(missing_index, m) for m in self.modality_types], dim=1)
        completed = self.complete_gcn(feats, _adjacency(~missing))
        # the missing modality's embedding is replaced by its reconstruction (:316-318, in place there)
        feats = torch.where(missing[:, :, None], completed, feats)
        full = torch.ones_like(missing)
        out = self.fusion_gcn(feats, _adjacency(full)).mean(dim=-2)
        return self.head(self.norm(out))


_FUSIONS = {
    'graph_fusion': modal_graph_fusion, 'unified_graph': modal_unified_graph,
    'sum': modal_sum, 'concat': modal_concat, 'regression': modal_regression, 'retrieval': modal_concat_full,
    'intra_attention': modal_intra_channel_attention, 'inter_attention': modal_inter_attention,
    'dedicated_dnn': modal_dedicated_dnn, 'Distill_tea': modal_distillation, 'MTD_stu': modal_distillation,
    'KL_stu': modal_distillation, 'self_distill': modal_self_distillation,
}


class finetune_model(nn.Module):
    def __init__(self, args, output_dims, encoder_model):
        super().__init__()
        self.encoder = encoder_model
        self.fusion_type = args.fusion_type
        if args.fusion_type not in _FUSIONS:
            raise ValueError(f"unknown fusion_type {args.fusion_type!r}")
        self.fusion = _FUSIONS[args.fusion_type](args, output_dims)

    def forward(self, data, missing_index):
        # host inputs: the bank uploads every tower's tensors on that tower's stream; missing_index goes up once
        if torch.is_tensor(missing_index) and not missing_index.is_cuda:
            p = next(self.parameters(), None)
            if p is not None and p.is_cuda:
                missing_index = missing_index.to(p.device, non_blocking=True)
        # `retrieval` ignores missing_index (the loader substituted a same-label sample and reset the
        # code to 0, data_loader.py:271-276), so nothing may be skipped for it
        # the graph heads read a missing modality's embedding through its self-loop: not skip-safe either
        skip_safe = self.fusion_type != 'retrieval' and getattr(self.fusion, 'compaction_safe', True)
        if getattr(self.encoder, 'supports_compaction', False) and skip_safe:
            embedding = self.encoder(data, missing_index=missing_index)   # towers skip missing samples
        else:
            embedding = self.encoder(data)
        return self.fusion(embedding, missing_index)
