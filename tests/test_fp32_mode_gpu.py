"""fp32 VERIFICATION mode on the B200 (MISSM_PRECISION=fp32 / autograd.set_precision('fp32')): the kernels of
csrc/fp32_mode.cu against their CPU statements (tests/ops_emulation.py, float64 where it matters), then the whole
path against the goldens of the unmodified reference at the north_star's fp32 tolerance: embeddings and loss
<= 1e-5 relative.  Run with `-m gpu`; /root/reference is not needed."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
import restatement as R  # noqa: E402
import ops_emulation as E  # noqa: E402  (CPU statements of the ABI contracts: the checker)

GOLD = os.path.join(HERE, "golden")
DEV = "cuda"
TOL_F32 = 1e-5        # north_star: embeddings / loss in fp32 mode
TOL_F32_GRAD = 1e-4   # gradients: fp32 summation order (split-K atomics, per-CTA partials) differs from the reference


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture()
def fp32_mode():
    from missm_b200 import autograd as ag
    old = ag.set_precision("fp32")
    yield
    ag.set_precision(old)


def test_expand6_bit_exact():
    from missm_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(37, 588) * torch.logspace(-3, 3, 588)
    xd = x.to(DEV)
    for which in (0, 1):
        assert torch.equal(ops.expand6(xd, which, False, 592).cpu(), E.expand6(x, which, False, 592))
        assert torch.equal(ops.expand6(xd, which, True).cpu(), E.expand6(x, which, True))
    v = xd[:, 8:72]                                                      # strided view (ld 588)
    assert torch.equal(ops.expand6(v, 0, False, 64).cpu(), E.expand6(x[:, 8:72].contiguous(), 0, False, 64))


@pytest.mark.parametrize("M,N,K", [(40, 64, 588), (136, 24, 72), (8, 8, 1021), (2056, 1024, 1024), (1504, 256, 4096)])
def test_split_gemm_fp32_accuracy(M, N, K):
    """One tcgen05 bf16 launch over the 3-way split operands vs float64: all operand layouts.
    Measured on B200: 9e-8 (K = 72), 4-9e-7 (K = 588 / 1021), 1.0e-6 (K = 1024), 2.9e-6 (K = 4096, split-K) and
    4.4e-6 (K = 4096, one accumulator): the error grows with the length of ONE accumulation chain, i.e. the
    tensor core's fp32 accumulate truncates rather than rounds -- still below the 1e-5 the mode promises."""
    from missm_b200 import ops
    torch.manual_seed(2)
    a, b = torch.randn(M, K), torch.randn(N, K) * 3
    ref = a.double() @ b.double().t()
    ad, bd = a.to(DEV), b.to(DEV)
    errs = [rel(ops.gemm_f32(ad, bd), ref)]
    if K % 8 == 0:
        errs.append(rel(ops.gemm_f32(ad, bd.t().contiguous(), b_mn=True), ref))
    errs.append(rel(ops.gemm_f32(ad.t().contiguous(), bd.t().contiguous(), a_mn=True, b_mn=True), ref))
    bias = torch.randn(N)
    res = torch.randn(M, N)
    errs.append(rel(ops.gemm_f32(ad, bd, bias=bias.to(DEV), epilogue=ops.EPI_RESID, aux_in=res.to(DEV)),
                    ref + bias.double() + res.double()))
    print('gemm_f32', (M, N, K), errs)
    assert max(errs) < 1e-6 + 1.5e-9 * K, errs
    assert rel(torch.matmul(ad.bfloat16(), bd.bfloat16().t()).float(), ref) > 1e-3     # what one bf16 piece gives


def _layouts():
    from missm_b200 import ops
    return {'spatial': (ops.SeqLayout.spatial(3, 50), 150), 'temporal': (ops.SeqLayout.temporal(2, 4, 17), 2 * 4 * 17)}


@pytest.mark.parametrize("kind", ['spatial', 'temporal', 'causal_mask'])
def test_attention_f32_kernels(kind):
    from missm_b200 import ops
    torch.manual_seed(3)
    H, D = 2, 128
    kw = {}
    if kind == 'causal_mask':
        lay, rows = ops.SeqLayout.spatial(3, 77), 231
        km = torch.ones(5, 77, dtype=torch.int64)
        km[1, 21:] = 0
        km[4, 9:] = 0
        kw = dict(causal=True, key_mask=km, mask_rows=torch.tensor([4, 0, 1], dtype=torch.int32))
    else:
        lay, rows = _layouts()[kind]
    qkv = torch.randn(rows, 3 * D)
    qkv[:, :D] *= 0.125
    d_out = torch.randn(rows, D)
    o_ref, lse_ref = E.attention_f32_fwd(qkv.double().float(), lay, H, **kw)
    g_ref = E.attention_f32_bwd(qkv, o_ref, lse_ref, d_out, lay, H, 0.125, **kw)
    kwd = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in kw.items()}
    o, lse = ops.attention_f32_fwd(qkv.to(DEV), lay, H, **kwd)
    g = ops.attention_f32_bwd(qkv.to(DEV), o, lse, d_out.to(DEV), lay, H, 0.125, **kwd)
    print(kind, rel(o, o_ref), rel(lse, lse_ref), rel(g, g_ref))
    assert rel(o, o_ref) < 2e-6 and rel(lse, lse_ref) < 2e-6 and rel(g, g_ref) < 5e-6


def test_small_fp32_kernels():
    from missm_b200 import ops
    torch.manual_seed(4)
    u, d = torch.randn(1000, 256) * 3, torch.randn(1000, 256)
    assert rel(ops.gelu_f32_fwd(u.to(DEV)), E.gelu_f32_fwd(u.double())) < 1e-6
    assert rel(ops.gelu_f32_bwd(d.to(DEV), u.to(DEV)), E.gelu_f32_bwd(d.double(), u.double())) < 1e-6
    assert rel(ops.colsum_f32(u.to(DEV)), u.double().sum(0)) < 1e-6
    px = torch.randn(5, 3, 4, 28, 42)
    idx = torch.tensor([4, 0, 2], dtype=torch.int32)
    got = ops.patchify_f32(px.to(DEV), 14, 592, 4, sample_index=idx.to(DEV), n_samples=3)
    assert torch.equal(got.cpu(), E.patchify_f32(px, 14, 592, 4, sample_index=idx, n_samples=3))


@pytest.fixture(scope="module")
def tiny():
    return torch.load(os.path.join(GOLD, "tiny_bank.pt"), weights_only=False)


def test_fp32_mode_matches_reference_golden(tiny, fp32_mode):
    """All five towers + text + `sum` head with mask compaction, fwd + bwd, vs the UNMODIFIED reference's fp32
    results (tests/golden/tiny_bank.pt): embeddings and loss <= 1e-5, gradients <= 1e-4."""
    from test_parity_gpu import make, to_dev
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    model, cfgs, tcfg, _ = make(meta, modal_types, 'sum')
    model.train()
    data = to_dev(R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0))
    mi = tiny['missing_index'].to(DEV)
    with torch.no_grad():
        emb = model.encoder(data)
    logits = model(data, mi)
    loss = torch.nn.functional.cross_entropy(logits, tiny['labels'].to(DEV))
    loss.backward()
    errs = {m: rel(emb[m], tiny[f'emb/{m}']) for m in modal_types}
    loss_err = abs(loss.item() - tiny['loss/sum'].item()) / abs(tiny['loss/sum'].item())
    print('fp32 mode: emb', errs, 'logits', rel(logits, tiny['logits/sum']), 'loss', loss_err)
    for m in modal_types:
        assert errs[m] < TOL_F32, (m, errs)
    assert rel(logits, tiny['logits/sum']) < 1e-4
    assert loss_err < TOL_F32
    params = dict(model.named_parameters())
    worst = ('', 0.0)
    for k, v in tiny.items():
        if k.startswith('grad/') and v.norm() > 1e-6:
            e = rel(params[k[5:]].grad, v)
            worst = max(worst, (k, e), key=lambda t: t[1])
    print('fp32 mode: worst gradient', worst)
    assert worst[1] < TOL_F32_GRAD, worst
    for n, ref in tiny['grad_norms'].items():
        if ref > 1e-6:
            assert abs(params[n].grad.norm().item() - ref) < 1e-4 * ref, n


def test_fp32_and_bf16_modes_share_parameters(tiny, fp32_mode):
    """Switching modes needs no reload: the bf16 operand caches and the fp32 packed weights live side by side."""
    from missm_b200 import autograd as ag
    from test_parity_gpu import make, to_dev
    meta = tiny['meta']
    model, cfgs, tcfg, _ = make(meta, ['image'], 'sum')
    model.eval()
    data = to_dev(R.synth_inputs(['image'], 4, cfgs, tcfg, seed=2))
    with torch.no_grad():
        e32 = model.encoder(data)['image']
        ag.set_precision("bf16")
        e16 = model.encoder(data)['image']
        ag.set_precision("fp32")
        e32b = model.encoder(data)['image']
    assert torch.equal(e32, e32b)
    assert 1e-5 < rel(e16, e32) < 1e-2
