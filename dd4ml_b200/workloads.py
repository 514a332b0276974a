"""Synthetic workloads of the BASELINE configs (bench / smoke support).

Models are plain PyTorch re-statements of the reference architectures (random init of
that architecture, synthetic data of that shape -- there is no network for datasets);
their forward/backward is untouched PyTorch, as the north star requires.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class MediumFFNN(nn.Module):
    """784 -> [128] x 8 -> 10 with log_softmax (reference models/ffnn/medium_ffnn.py:6-39,
    tests/config_files/config_apts_d.yaml:48-49): n = 217 354 parameters, 18 tensors."""

    def __init__(self, in_features=784, hidden=(128,) * 8, classes=10):
        super().__init__()
        dims = [in_features, *hidden]
        self.stages = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.finish = nn.Linear(dims[-1], classes)

    def forward(self, x):
        x = x.reshape(x.size(0), -1)
        for lin in self.stages:
            x = F.relu(lin(x))
        return F.log_softmax(self.finish(x), dim=1)


def mnist_shape_batch(n_samples, seed, device="cpu", pin=False):
    """x ~ N(0,1) of shape (B, 1, 28, 28), y ~ U{0..9}."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_samples, 1, 28, 28, generator=g)
    y = torch.randint(0, 10, (n_samples,), generator=g)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return x.to(device), y.to(device)


class Cfg:
    """Attribute bag with the keys of the reference's W&B/yaml config contract."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def apts_d_yaml_config(**over):
    """tests/config_files/config_apts_d.yaml:6-83 (the configuration BASELINE.json's
    configs[1] is quoted on)."""
    c = dict(delta=0.1, min_delta=0.01, max_delta=1.0, norm_type=2, glob_second_order=True,
             glob_dogleg=True, loc_second_order=True, loc_dogleg=True, mem_length=3, max_wolfe_iters=5,
             max_zoom_iters=5, tol=1e-6, paper_tr_update=False, glob_opt="lssr1_tr", loc_opt="lssr1_tr",
             max_loc_iters=3, max_glob_iters=1, glob_pass=True, foc=True)
    c.update(over)
    return Cfg(**c)


class PINNNet(nn.Module):
    """1 -> [20] x 8 -> 1 with tanh (reference models/ffnn/pinn_ffnn.py:11-29): n = 3 001 parameters,
    18 tensors in the same order (weight, bias per layer)."""

    def __init__(self, in_features=1, hidden=(20,) * 8, out_features=1):
        super().__init__()
        dims = [in_features, *hidden]
        mods = []
        for a, b in zip(dims[:-1], dims[1:]):
            mods += [nn.Linear(a, b), nn.Tanh()]
        mods.append(nn.Linear(dims[-1], out_features))
        self.net = nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)


class AllenCahnLoss(nn.Module):
    """Residual loss of the 1-D Allen-Cahn PINN, -u'' + u^3 - u = 0 on [low, high], u(low) = 1,
    u(high) = -1 (reference utility/pinn_allencahn_loss.py:14-60): mean over ALL points of the squared
    interior residual masked by (1 - flag) plus mean over all points of the squared boundary error
    masked by flag.  ``current_x`` is the collocation tensor the prediction was computed from (set by
    the caller: trainer.py:997-998, apts_pinn.py:_set_criterion_input)."""

    def __init__(self, current_x=None, low=0.0, high=1.0):
        super().__init__()
        self.current_x, self.low, self.high = current_x, low, high

    def forward(self, u, flag):
        x = self.current_x
        if x is None:
            raise ValueError("current_x must be set before calling AllenCahnLoss")
        if not x.requires_grad:
            x.requires_grad_(True)
        ones = torch.ones_like(u)
        du, = torch.autograd.grad(u, x, ones, create_graph=True, retain_graph=True)
        d2u, = torch.autograd.grad(du, x, torch.ones_like(du), create_graph=True)
        flag = flag.squeeze()
        xs, us = x.squeeze(), u.squeeze()
        res = (-d2u + u ** 3 - u).squeeze()
        target = torch.zeros_like(xs)
        target[torch.isclose(xs, torch.tensor(self.low, device=x.device, dtype=x.dtype))] = 1.0
        target[torch.isclose(xs, torch.tensor(self.high, device=x.device, dtype=x.dtype))] = -1.0
        return (res ** 2 * (1 - flag)).mean() + ((us - target) ** 2 * flag).mean()


def allen_cahn_points(n_interior=1000, low=0.0, high=1.0):
    """Collocation set of datasets/pinn_allencahn.py:38-64: n_interior equispaced interior points plus
    the two boundary points, sorted; returns (x [N,1] fp32, boundary flag [N,1] fp32)."""
    step = (high - low) / (n_interior + 1)
    interior = torch.linspace(low + step, high - step, n_interior, dtype=torch.float32).unsqueeze(1)
    boundary = torch.tensor([[low], [high]], dtype=torch.float32)
    data = torch.cat([boundary, interior], 0)
    mask = torch.cat([torch.ones(2, 1), torch.zeros(n_interior, 1)], 0)
    idx = torch.argsort(data[:, 0])
    return data[idx], mask[idx]
