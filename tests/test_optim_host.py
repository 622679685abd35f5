"""CPU tests of the host side of missm_b200.optim.FusedAdam (reference call sites
train_ddp.py:205,221,254): constructor contract of torch.optim.Adam, chunk table, bias-correction scalars, state
layout, the opt-in rebinding of torch.optim.Adam, and the loud failure without a CUDA device.  The kernel itself
(csrc/optim.cu) is checked against torch.optim.Adam on the B200 by tests/test_optim_gpu.py."""
import math
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_chunk_table_covers_every_element_once():
    from missm_b200 import optim
    numels = [5, 70000, 0, 32768, 32769]
    t, off = optim.chunk_table(numels, 32768)
    covered = [0] * len(numels)
    for ti, o in zip(t, off):
        assert o % 32768 == 0 and o < numels[ti]
        covered[ti] += min(32768, numels[ti] - o)
    assert covered == numels and len(t) == 1 + 3 + 0 + 1 + 2


def test_step_scalars_are_torch_adams():
    from missm_b200 import optim
    for step in (1, 2, 10, 1000):
        s, b = optim.step_scalars(step, 1e-4, 0.9, 0.999)
        assert s == 1e-4 / (1 - 0.9 ** step) and b == math.sqrt(1 - 0.999 ** step)


def test_constructor_contract_and_state_layout():
    from missm_b200 import optim
    p = torch.nn.Parameter(torch.zeros(3))
    o = optim.FusedAdam([p], lr=1e-4, weight_decay=0.0)           # the call of train_ddp.py:205
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3))], lr=1e-4, weight_decay=0.0)
    for k in ('lr', 'betas', 'eps', 'weight_decay', 'amsgrad'):
        assert o.param_groups[0][k] == ref.param_groups[0][k], k
    for bad in (dict(lr=-1.0), dict(eps=-1.0), dict(betas=(1.0, 0.9)), dict(betas=(0.9, 1.0)), dict(weight_decay=-1)):
        with pytest.raises(ValueError):
            optim.FusedAdam([p], **bad)
    with pytest.raises(NotImplementedError):
        optim.FusedAdam([p], amsgrad=True)
    st = o._init_state(p)
    assert set(st) == {'step', 'exp_avg', 'exp_avg_sq'} and st['step'].item() == 0.0
    # ReduceLROnPlateau (train_ddp.py:206) drives it through param_groups
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(o, mode='max', factor=0.1, patience=0)
    sched.step(1.0), sched.step(0.5)
    assert abs(o.param_groups[0]['lr'] - 1e-5) < 1e-12


def test_no_cpu_fallback():
    from missm_b200 import optim
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        optim.FusedAdam([p]).step()


def test_install_rebinds_and_restores():
    from missm_b200 import optim
    orig = torch.optim.Adam
    optim.install()
    try:
        assert torch.optim.Adam is optim.FusedAdam
        from torch import optim as script_optim                  # how train_ddp.py spells it
        assert script_optim.Adam is optim.FusedAdam
    finally:
        optim.uninstall()
    assert torch.optim.Adam is orig


def test_operand_refresh_plan_only_touches_exact_entries():
    """autograd.plan_operand_refresh / restamp: which cached bf16 operand copies an optimizer may rewrite itself."""
    import ops_emulation as E
    from missm_b200 import autograd as ag
    D = 16
    mk = lambda *s: torch.nn.Parameter(torch.randn(*s))
    qw, kw, vw, ow = mk(D, D), mk(D, D), mk(D, D), mk(D, D)
    qb, kb, vb = mk(D), mk(D), mk(D)
    patch = mk(D, 3, 2, 2)
    cache = {}
    with E.emulated_fp32_mode(precision="bf16"):
        w, b = ag.packed_qkv(cache, qw, kw, vw, qb, kb, vb)
        wo = ag.bf16_weight(cache, "o", ow)
        ag.bf16_weight(cache, "patch", patch, cols_dst=16)          # padded layout: never a sink
        # everything fresh, all weights touched: q/k/v slices of the pack + o qualify, the padded patch copy does not
        dst, entries = ag.plan_operand_refresh([qw, kw, vw, ow, qb, patch])
        assert set(dst) == {id(qw), id(kw), id(vw), id(ow)}
        assert dst[id(kw)].data_ptr() == w[D:2 * D].data_ptr() and dst[id(ow)].data_ptr() == wo.data_ptr()
        assert sorted(n for _, n in entries) == ["o", "qkv_w"]
        # simulate the optimizer: raw-pointer update = new values + version bump, bf16 slices rewritten
        with torch.no_grad():
            for p_ in (qw, kw, vw, ow, qb):
                p_.add_(1.0)
            for p_ in (qw, kw, vw, ow):
                dst[id(p_)].copy_(p_.detach().to(torch.bfloat16))
        ag.restamp(entries)
        w2, b2 = ag.packed_qkv(cache, qw, kw, vw, qb, kb, vb)
        assert w2.data_ptr() == w.data_ptr() and torch.equal(w2[:D], qw.detach().to(torch.bfloat16))   # not rebuilt
        assert b2.data_ptr() != b.data_ptr() and torch.equal(b2[:D], qb.detach())                      # biases rebuilt
        # an entry that is already stale (k changed behind the cache's back) is left to rebuild itself
        with torch.no_grad():
            kw.mul_(2.0)
        dst, entries = ag.plan_operand_refresh([qw, vw])
        assert dst == {} and entries == []
        w3, _ = ag.packed_qkv(cache, qw, kw, vw, qb, kb, vb)
        assert torch.equal(w3[D:2 * D], kw.detach().to(torch.bfloat16))
