"""CPU tests of the fp32 VERIFICATION mode's host side (missm_b200/autograd_f32.py, ops.gemm_f32): with the C-ABI
entry points replaced by torch stand-ins written from the ABI contracts (tests/ops_emulation.py -- test
infrastructure, the product has no CPU path) the unchanged host code must reproduce the reference's goldens at
fp32 accuracy.  This pins, without a GPU: the 3-way bf16 split (piece order, padding, K-major / MN-major
stacking of the expanded operands), every gradient formula of the fp32 blocks, and the strided temporal layout.
The CUDA kernels behind the same calls are checked on the B200 by tests/test_fp32_mode_gpu.py."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
import restatement as R  # noqa: E402
import ops_emulation as E  # noqa: E402

GOLD = os.path.join(HERE, "golden")
TOL_F32 = 1e-5       # north_star: embeddings / loss <= 1e-5 relative in fp32 mode
TOL_F32_GRAD = 1e-4  # gradients (through up to 2 layers x 6 towers, fp32 summation order differs)


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_split_gemm_reaches_fp32_accuracy():
    """ops.gemm_f32 (real host code) over emulated expand6 + bf16 GEMM: all four operand layouts, ragged K."""
    from missm_b200 import ops
    torch.manual_seed(0)
    with E.emulated_fp32_mode():
        for (M, N, K) in ((40, 64, 588), (136, 24, 72), (8, 8, 1021)):
            a, b = torch.randn(M, K), torch.randn(N, K) * 3
            ref = a.double() @ b.double().t()
            assert rel(ops.gemm_f32(a, b), ref) < 3e-7                                   # K-major x K-major
            if K % 8 == 0:
                assert rel(ops.gemm_f32(a, b.t().contiguous(), b_mn=True), ref) < 3e-7   # dgrad layout
            assert rel(ops.gemm_f32(a.t().contiguous(), b.t().contiguous(), a_mn=True, b_mn=True), ref) < 3e-7
        # one bf16 piece alone is ~3 decimal digits: the split is what buys the accuracy
        a, b = torch.randn(16, 256), torch.randn(8, 256)
        assert rel((a.bfloat16().double() @ b.bfloat16().double().t()), a.double() @ b.double().t()) > 1e-3


@pytest.fixture(scope="module")
def tiny():
    return torch.load(os.path.join(GOLD, "tiny_bank.pt"), weights_only=False)


def _make(meta, modal_types, fusion):
    from missm_b200 import shapes
    per = meta.get('per', {})
    cfgs = {}
    for m in modal_types:
        if m == 'language':
            continue
        d = {k: v for k, v in meta['vision'].items() if k != 'lora_r'}
        d.update(per.get(m, {}))
        d['temporal_mlp'] = (m != 'video')
        cfgs[m] = R.vision_config(**d)
    tcfg = R.text_config(**meta['text'])
    model = shapes.build_finetune(cfgs, tcfg, modal_types, fusion, 3, meta.get('projection_dim', 768),
                                  meta.get('fusion_dim', 256), dropout_prob=0.0)
    sd = R.synth_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()])
    shapes.load_named(model, sd)
    return model, cfgs, tcfg


def test_fp32_mode_host_wiring_matches_reference_golden(tiny):
    """All five towers + text + `sum` head, one sample missing per modality: embeddings, loss and the stored
    gradients of the UNMODIFIED reference (tests/golden/tiny_bank.pt) at fp32 tolerances."""
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    model, cfgs, tcfg = _make(meta, modal_types, 'sum')
    model.train()
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0)
    mi = tiny['missing_index']
    with E.emulated_fp32_mode():
        with torch.no_grad():
            emb = model.encoder(data)
        logits = model(data, mi)
        loss = torch.nn.functional.cross_entropy(logits, tiny['labels'])
        loss.backward()
    for m in modal_types:
        assert rel(emb[m], tiny[f'emb/{m}']) < TOL_F32, (m, rel(emb[m], tiny[f'emb/{m}']))
    assert rel(logits, tiny['logits/sum']) < 1e-4        # differences of O(1) features in a random head
    assert abs(loss.item() - tiny['loss/sum'].item()) < TOL_F32 * abs(tiny['loss/sum'].item())
    params = dict(model.named_parameters())
    worst = 0.0
    for k, v in tiny.items():
        if k.startswith('grad/') and v.norm() > 1e-6:
            e = rel(params[k[5:]].grad, v)
            worst = max(worst, e)
            assert e < TOL_F32_GRAD, (k, e)
    for n, ref in tiny['grad_norms'].items():
        if ref > 1e-6:
            assert abs(params[n].grad.norm().item() - ref) < 1e-4 * ref, n


def test_lockstep_issue_order_does_not_change_results(tiny):
    """bank.LanguageBind.forward issues the towers one encoder layer at a time in turn (lockstep) or tower by tower
    (MISSM_LOCKSTEP=0): same kernels, same operands, another order -- outputs and gradients must be identical."""
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    model, cfgs, tcfg = _make(meta, modal_types, 'sum')
    model.train()
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0)
    mi = tiny['missing_index']
    res = {}
    with E.emulated_fp32_mode():
        for mode in (True, False):
            model.encoder.lockstep = mode
            model.zero_grad(set_to_none=True)
            logits = model(data, mi)
            torch.nn.functional.cross_entropy(logits, tiny['labels']).backward()
            res[mode] = (logits.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    assert torch.equal(res[True][0], res[False][0])
    assert res[True][1].keys() == res[False][1].keys()
    for n in res[True][1]:
        assert torch.equal(res[True][1][n], res[False][1][n]), n


def test_precision_switch():
    from missm_b200 import autograd as ag
    assert ag.get_precision() == os.environ.get("MISSM_PRECISION", "bf16")
    old = ag.set_precision("fp32")
    try:
        assert ag.get_precision() == "fp32"
        with pytest.raises(ValueError):
            ag.set_precision("fp16")
    finally:
        ag.set_precision(old)


def test_product_blocks_host_algebra_matches_reference_golden(tiny):
    """The PRODUCT (bf16-mode) autograd Functions -- AttnBlockFn, MlpBlockFn, VisionEmbedFn, TextEmbedFn, PoolProjFn,
    ScatterZeroFn with mask compaction, the side channel of bf16 gradient copies and fused bias sums -- over the ABI
    stand-ins with every "bf16" buffer widened to fp32: all five towers + text + `sum` head must reproduce the
    reference's embeddings, loss and gradients at fp32 accuracy, i.e. every formula and operand layout of the path
    the B200 runs is right independently of bf16 rounding (which the GPU tests bound at 1e-2)."""
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    model, cfgs, tcfg = _make(meta, modal_types, 'sum')
    model.train()
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0)
    mi = tiny['missing_index']
    with E.emulated_fp32_mode(precision="bf16", wide_bf16=True):
        with torch.no_grad():
            emb = model.encoder(data)
        logits = model(data, mi)
        loss = torch.nn.functional.cross_entropy(logits, tiny['labels'])
        loss.backward()
    for m in modal_types:
        assert rel(emb[m], tiny[f'emb/{m}']) < TOL_F32, (m, rel(emb[m], tiny[f'emb/{m}']))
    assert rel(logits, tiny['logits/sum']) < 1e-4
    assert abs(loss.item() - tiny['loss/sum'].item()) < TOL_F32 * abs(tiny['loss/sum'].item())
    params = dict(model.named_parameters())
    for k, v in tiny.items():
        if k.startswith('grad/') and v.norm() > 1e-6:
            assert rel(params[k[5:]].grad, v) < TOL_F32_GRAD, (k, rel(params[k[5:]].grad, v))
    for n, ref in tiny['grad_norms'].items():
        if ref > 1e-6:
            assert abs(params[n].grad.norm().item() - ref) < 2e-4 * ref, n
