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


def sigma_spread(info, delta, gamma=None, dtype=torch.float32):
    """How much the reference's OBS result depends on the (implementation-defined) signs LAPACK gives the
    eigenvectors: relative spread of sigma* over all 2^k sign patterns of a_1..a_k (obs.py:162-170 starts
    Newton from max_j(a_j/delta - Lambda_j) with SIGNED a_j; the last a is a norm).  0.0 = sign-insensitive.
    `info` is the dict filled by oracle.obs_solve (Lambda, a, case)."""
    import itertools
    from oracle.obs import OBS_TOL, _newton
    Lam, a = info["Lambda"].to(dtype), info["a"].to(dtype)
    k = a.numel() - 1
    lam_min = float(info["lam_min"]) if "lam_min" in info else min(float(Lam[0]), float(gamma))
    if info["case"] != "C" or lam_min > 0:
        return 0.0
    d = torch.tensor(delta, dtype=dtype)
    sig = []
    for s in itertools.product((1.0, -1.0), repeat=k):
        aa = a.clone()
        aa[:k] = aa[:k] * torch.tensor(s, dtype=dtype)
        sh = torch.max(aa / d - Lam)
        x0 = sh if sh > -lam_min else torch.tensor(-lam_min, dtype=dtype)
        sig.append(float(_newton(x0, Lam, aa, d, OBS_TOL)[0]))
    return (max(sig) - min(sig)) / max(abs(max(sig)), 1e-30)


def golden_solves(name):
    """The sub-problem solves recorded inside a trajectory golden (gen_golden.SolveRecorder): dicts with
    g, S, Y (n x k), gamma, delta, dogleg, p, pred of the UNMODIFIED reference."""
    g = load_golden(name)
    out = []
    for i in range(int(g["n_solves"])):
        out.append({k: g[f"solve{i}_{k}"] for k in ("g", "S", "Y", "gamma", "delta", "dogleg", "p", "pred", "tol")})
    return out


def oracle_memory(S, Y, gamma, dtype=torch.float32):
    """LSR1Memory holding exactly the given pairs / gamma (bypasses the acceptance chain)."""
    from oracle.lsr1 import LSR1Memory
    S, Y = torch.as_tensor(S).to(dtype), torch.as_tensor(Y).to(dtype)
    h = LSR1Memory(gamma=float(gamma), memory_length=max(S.shape[1], 1), tol=1e-6, dtype=dtype)
    h.S = [S[:, j].clone() for j in range(S.shape[1])]
    h.Y = [Y[:, j].clone() for j in range(Y.shape[1])]
    h.gamma = torch.tensor(float(gamma), dtype=dtype)
    return h
