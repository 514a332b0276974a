"""APTS_PINN (spatial subdomains, apts_pinn.py) on one GPU against the oracle: the subdomain
mask, the n_local/N weighting of the local reduction and the full/sub closures."""
import math

import numpy as np
import pytest
import torch

from _helpers import FFNN, load_golden, make_fg, rel_err, set_flat, sine_crit

pytestmark = pytest.mark.gpu

import oracle

DEV = "cuda"


class SineCrit(torch.nn.Module):
    low, high = 0.0, 1.0

    def __init__(self):
        super().__init__()
        self.current_x = None

    def forward(self, out, labels):
        return ((out.squeeze(-1) - torch.sin(2 * math.pi * self.current_x.squeeze(-1))) ** 2).mean()


HP = dict(delta=0.1, nu_dec=0.25, nu_inc=0.75, inc_factor=1.2, dec_factor=0.9, max_delta=2.0,
          min_delta=1e-3, tol=1e-6)


def test_apts_pinn_single_rank_matches_oracle():
    from dd4ml_b200 import APTS_PINN, TR
    g = load_golden("apts_pinn_ws2")
    x = torch.from_numpy(g["x"])
    model = FFNN(1, [16, 16], 1, log_softmax=False)
    set_flat(model, torch.from_numpy(g["w0"]))
    model = model.to(DEV)
    opt = APTS_PINN(model.parameters(), model=model, criterion=SineCrit(), device=DEV, glob_opt=TR,
                    glob_opt_hparams=HP, loc_opt=TR, loc_opt_hparams=HP, glob_pass=True, foc=True, norm_type=2,
                    max_loc_iters=4, max_glob_iters=1, num_subdomains=1, overlap=0.1, **HP)
    assert opt.get_subdomain_mask(x[:, 0].to(DEV), 0).all()
    ofg = [make_fg(FFNN(1, [16, 16], 1, log_softmax=False), x, None, crit=sine_crit(x))]
    oopt = oracle.APTSDOracle(torch.from_numpy(g["w0"]), 1, "tr", HP, "tr", HP,
                              **{k: v for k, v in HP.items() if k != "tol"}, glob_pass=True, foc=True,
                              norm_type=2, max_loc_iters=4, max_glob_iters=1, tol=1e-6)
    xd, lab = x.to(DEV), torch.zeros(x.shape[0], 1, device=DEV)
    losses, olosses = [], []
    for _ in range(8):
        losses.append(float(opt.step(inputs=xd, labels=lab)))
        olosses.append(float(oopt.step(ofg, weights=[1.0])))
    np.testing.assert_allclose(losses, olosses, rtol=2e-3)
    assert abs(opt.delta - oopt.delta) < 1e-9


def test_apts_pinn_masks_overlap_and_cover():
    from dd4ml_b200 import APTS_PINN, TR
    model = FFNN(1, [8], 1, log_softmax=False).to(DEV)
    opt = APTS_PINN(model.parameters(), model=model, criterion=SineCrit(), device=DEV, glob_opt=TR,
                    glob_opt_hparams=HP, loc_opt=TR, loc_opt_hparams=HP, glob_pass=False, foc=False, norm_type=2,
                    max_loc_iters=1, max_glob_iters=1, num_subdomains=2, overlap=0.2, **HP)
    x = torch.linspace(0, 1, 12)
    m0, m1 = opt.get_subdomain_mask(x, 0), opt.get_subdomain_mask(x, 1)
    assert (m0 & m1).any() and (m0 | m1).all()        # reference tests/test_pinn_subdomains.py:86-138
