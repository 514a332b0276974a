"""GPU parity of TRAdam against the reference goldens and the CPU oracle."""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn

from _helpers import FFNN, load_golden, make_fg, rel_err, set_flat

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from dd4ml_b200 import TRAdam

import oracle

DEV = "cuda"
IN_F, HID, OUT = 36, [12, 12], 10


def flat_of(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()


@pytest.mark.parametrize("name,norm,lr", [("tradam_inf", math.inf, 0.01), ("tradam_l2", 2, 0.5)])
def test_tradam_matches_reference_golden(name, norm, lr):
    g = load_golden(name)
    X, Y = torch.from_numpy(g["X"]).to(DEV), torch.from_numpy(g["Y"]).to(DEV)
    model = FFNN(IN_F, HID, OUT)
    set_flat(model, torch.from_numpy(g["w0"]))
    model = model.to(DEV)
    opt = TRAdam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, norm_type=norm)
    crit = nn.CrossEntropyLoss()

    def closure():
        opt.zero_grad()
        loss = crit(model(X), Y)
        loss.backward()
        return loss

    syncs0 = opt._ds.syncs
    losses = []
    for i in range(int(g["steps"])):
        losses.append(opt.step(closure).detach())
        if i == 0:
            # t = 1: s = g / (|g| + eps): the step depends on the sign pattern of the gradient only
            assert rel_err(flat_of(model), g["w_step1"]) < 1e-6
    assert opt._ds.syncs == syncs0                      # no host read inside step()
    losses = [float(x) for x in losses]
    np.testing.assert_allclose(losses, g["loss"], rtol=2e-4)
    assert rel_err(flat_of(model), g["w_final"]) < 2e-4
    assert rel_err(opt._m_buf.cpu(), g["m_final"]) < 1e-3
    assert rel_err(opt._v_buf.cpu(), g["v_final"]) < 1e-3
    assert float(opt._step_buf.abs().max()) == pytest.approx(float(g["step_max"][-1]), rel=1e-3)


@pytest.mark.parametrize("norm,lr,dtype", [(math.inf, 0.01, torch.float32), (2, 0.5, torch.float32),
                                           (math.inf, 0.003, torch.float64)])
@pytest.mark.parametrize("n", [1, 777, 65536 + 3, (1 << 20) + 5])
def test_tradam_kernels_match_oracle_on_identical_gradients(norm, lr, dtype, n):
    """Same gradient sequence on both sides: m, v, step and w agree to the last bits (the element
    arithmetic follows the reference's fp32 operation order, csrc/k_tradam.cu)."""
    gen = torch.Generator().manual_seed(n)
    w0 = torch.randn(n, generator=gen).to(dtype)
    grads = [torch.randn(n, generator=gen) * (10.0 ** (-i)) for i in range(4)]
    p = nn.Parameter(w0.clone().to(DEV))
    opt = TRAdam([p], lr=lr, norm_type=norm)
    o = oracle.TRAdamOracle(n, lr=lr, norm_type=norm)
    w = w0.clone()
    for gi in grads:
        p.grad = gi.to(DEV, dtype)
        opt.step()
        o.step(lambda _w, gi=gi: (torch.zeros(()), gi), w)
        assert opt.step_length == pytest.approx(o.step_len, rel=2e-6)
        assert rel_err(opt._m_buf.cpu(), o.m) < 2e-7
        assert rel_err(opt._v_buf.cpu(), o.v) < 2e-7
        assert rel_err(opt._step_buf.cpu(), o.step_buf) < 4e-6
        assert rel_err(p.detach().cpu(), w) < 2e-7
    assert opt.t == 4


def test_tradam_resume_and_reset():
    g = load_golden("tradam_inf")
    X, Y = torch.from_numpy(g["X"]).to(DEV), torch.from_numpy(g["Y"]).to(DEV)

    def build():
        model = FFNN(IN_F, HID, OUT)
        set_flat(model, torch.from_numpy(g["w0"]))
        model = model.to(DEV)
        opt = TRAdam(model.parameters(), lr=0.01)
        crit = nn.CrossEntropyLoss()

        def closure():
            opt.zero_grad()
            loss = crit(model(X), Y)
            loss.backward()
            return loss
        return model, opt, closure

    model, opt, closure = build()
    for _ in range(3):
        opt.step(closure)
    ck_m, ck_o = {k: v.clone() for k, v in model.state_dict().items()}, opt.state_dict()
    cont = [float(opt.step(closure).detach()) for _ in range(3)]
    model2, opt2, closure2 = build()
    model2.load_state_dict(ck_m)
    opt2.load_state_dict(ck_o)
    assert [float(opt2.step(closure2).detach()) for _ in range(3)] == cont
    opt2.reset_momentum()
    assert float(opt2._m_buf.abs().max()) == 0.0 and float(opt2._v_buf.abs().max()) == 0.0
    with pytest.raises(ValueError):
        TRAdam(model.parameters(), lr=0.1, norm_type=1)
