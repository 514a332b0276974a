"""Shared test helpers: tiny functional models over a flat parameter vector."""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


class FFNN(nn.Module):
    """Linear/ReLU stack + log_softmax: same architecture and parameter order as the
    reference's MediumFFNN (models/ffnn/medium_ffnn.py:17-39), written independently."""

    def __init__(self, in_f, hidden, out_f, log_softmax=True):
        super().__init__()
        dims = [in_f] + list(hidden)
        self.hidden = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.head = nn.Linear(dims[-1], out_f)
        self.log_softmax = log_softmax

    def forward(self, x):
        x = x.reshape(x.size(0), -1)
        for lin in self.hidden:
            x = F.relu(lin(x))
        x = self.head(x)
        return F.log_softmax(x, dim=1) if self.log_softmax else x


def set_flat(model, w):
    off = 0
    with torch.no_grad():
        for p in model.parameters():
            n = p.numel()
            p.copy_(w[off:off + n].view_as(p))
            off += n


def get_flat_grad(model):
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()


def make_fg(model, X, Y, crit=None):
    """fg(w) -> (loss, grad) evaluated by loading the flat vector into `model`."""
    crit = crit or nn.CrossEntropyLoss()

    def fg(w):
        set_flat(model, w)
        for p in model.parameters():
            p.grad = None
        loss = crit(model(X), Y)
        loss.backward()
        return loss.detach(), get_flat_grad(model)

    return fg


def sine_crit(x):
    def crit(out, _labels=None):
        return ((out.squeeze(-1) - torch.sin(2 * math.pi * x.squeeze(-1))) ** 2).mean()
    return crit


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def noisy_fg(fg, eps, gen):
    """fg with every evaluation perturbed by a relative `eps` (size of GPU-vs-CPU rounding of the model)."""
    def f(w):
        l, g = fg(w)
        if eps > 0:
            g = g * (1 + eps * torch.randn(g.shape, generator=gen, dtype=g.dtype))
            l = l * (1 + eps * float(torch.randn((), generator=gen)))
        return l, g
    return f


def branch_distance(gpu_losses, oracle_run, k, seeds=10, eps=1e-6):
    """Second-order TR trajectories branch on discrete decisions (L-SR1 pair acceptance, Wolfe/zoom
    exits, accept/reject) that flip under 1e-6 perturbations; the oracle itself reaches a handful of
    distinct branches.  Returns the smallest max-relative distance over the first k losses between
    the GPU run and the oracle branches reached with relative-eps perturbed evaluations.
    oracle_run(eps, seed) -> list of losses."""
    runs = [oracle_run(0.0, 0)] + [oracle_run(eps, s) for s in range(seeds)]
    gpu = np.asarray(gpu_losses[:k], dtype=np.float64)
    d = [float(np.max(np.abs(gpu - np.asarray(r[:k])) / np.abs(np.asarray(r[:k])))) for r in runs]
    return min(d), runs[int(np.argmin(d))]
