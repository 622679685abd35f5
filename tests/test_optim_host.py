"""CPU tests of the host side of missm_b200.optim.FusedAdam (reference call sites
train_ddp.py:205,221,254): constructor contract of torch.optim.Adam, chunk table, bias-correction scalars, state
layout, the opt-in rebinding of torch.optim.Adam, and the loud failure without a CUDA device.  The kernel itself
(csrc/optim.cu) is checked against torch.optim.Adam on the B200 by tests/test_optim_gpu.py."""
import math

import pytest
import torch


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
