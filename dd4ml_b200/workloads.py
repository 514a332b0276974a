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
