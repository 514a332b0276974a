"""Oracle: first/second-order trust-region sub-problem wrappers
(reference ``utility/optimizer_utils.py:197-307``).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch

from .lsr1 import LSR1Memory
from .obs import obs_solve


def solve_first_order(g, gn, delta, tol):
    """optimizer_utils.py:197-215.  ``gn`` is the p-norm the caller chose
    (for p=inf the predicted reduction delta*gn is *not* -g^T p)."""
    if gn <= tol:
        return torch.zeros_like(g), 0.0
    step = -g * (delta / gn)
    return step, delta * gn


def solve_second_order(g, gn, delta, hess: LSR1Memory, tol, dogleg=False, *,
                       eig_sign="lapack", info=None):
    """optimizer_utils.py:218-307: OBS step, Cauchy point (always computed),
    optional dogleg, pred = -(g^T p + 1/2 p^T B p)."""
    if gn <= tol:
        return torch.zeros_like(g), 0.0
    hess.precompute()                                            # :247
    delta_t = torch.scalar_tensor(delta, dtype=g.dtype)
    p_b = obs_solve(g, delta_t, hess.gamma, hess.Psi, hess.Minv, eig_sign=eig_sign, info=info)
    Bg = hess.B(g).to(g.dtype)
    gBg = g.dot(Bg)
    alpha = gn ** 2 / gBg if gBg > 0 else 0.0                    # :265
    p_cand = -g * alpha
    p_c = -g * (delta / gn) if torch.norm(p_cand) >= delta else p_cand
    if dogleg and gBg > 0:                                       # :274
        if torch.norm(p_b) <= delta:
            p = p_b
        elif torch.norm(p_c) >= delta:
            p = p_c
        else:
            d = p_b - p_c
            a = d.dot(d)
            b = 2 * p_c.dot(d)
            c = p_c.dot(p_c) - delta ** 2
            disc = b ** 2 - 4 * a * c
            if disc < 0:
                p = p_c
            else:
                tau = (-b + torch.sqrt(torch.clamp(disc, min=0.0))) / (2 * a)
                p = p_c + tau * d
    else:
        p = p_b
    g_dot_p = g.dot(p.to(g.dtype))
    Bp = hess.B(p).to(p.dtype)
    pred = -(g_dot_p + 0.5 * p.dot(Bp))
    return p, pred
