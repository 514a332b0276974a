"""OBS (orthonormal-basis SR1) trust-region sub-problem solver, drop-in name for
``dd4ml.solvers.obs.OBS`` (obs.py:26-254).

On the hot path the OBS computation is not a separate call: it runs in k-space inside the
last CTA of ``tr_propose`` / ``lssr1_begin`` (csrc/kspace.h ``tr_solve``), fed by the Gram
blocks of the L-SR1 pairs.  This class keeps the reference's entry point for callers that
hold an ``LSR1`` object: ``solve_tr_subproblem(g, delta, gamma, Psi, Minv)`` accepts the
owning ``LSR1`` in place of ``Psi`` (its Psi / Minv are implicit) and returns p*.
"""
from __future__ import annotations

import torch


class OBS:
    def __init__(self):
        self.tol = 1e-6          # obs.py:29
        self.last: dict = {}     # diagnostics of the last solve (case, sigma*, Newton iterations)

    def solve_tr_subproblem(self, g, delta, gamma, Psi, Minv=None):
        from ..optimizers.lsr1 import LSR1
        from ..utility.optimizer_utils import solve_tr_second_order
        if not isinstance(Psi, LSR1):
            raise TypeError(
                "dd4ml_b200.OBS works on the device-resident L-SR1 state: pass the LSR1 object as `Psi` "
                "(explicit (n x k) Psi / Minv matrices are never materialised in this implementation)")
        h = Psi
        for name, t in (("g", g), ("delta", torch.as_tensor(delta)), ("gamma", torch.as_tensor(gamma))):
            if torch.isnan(t).any() or torch.isinf(t).any():
                raise ValueError(f"{name} contains NaN or Inf values")
        if float(delta) < 0:
            raise ValueError("Delta must be non-negative")
        dog = h.ds.read().dogleg
        p, _ = solve_tr_second_order(g, torch.norm(g), float(delta), h, self, h.tol, dogleg=False)
        h.ds.set(dogleg=dog)
        return p
