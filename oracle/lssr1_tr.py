"""Oracle: L-S-SR1 trust-region step with strong-Wolfe line search
(reference ``optimizers/lssr1_tr.py:280-654``).  TEST INFRASTRUCTURE ONLY.

Flat-vector restatement: ``w`` is the flat parameter vector (modified in place),
``fg(w) -> (loss, grad)``.  Quirks kept on purpose:
Q5 the momentum term is dead (old_wk is overwritten before it is used),
Q6 rho is 0 unless pred_red < 0, the step is never rejected.
The cross-rank ``sync`` of (loss, ||g||) and the gradient broadcast
(lssr1_tr.py:503-507, 271-276) are identities when every rank already holds the
same values, which is the case on the APTS path; an ``avg`` hook is accepted
for completeness.
"""
from __future__ import annotations

import math

import torch

from .lsr1 import LSR1Memory
from .subproblem import solve_first_order, solve_second_order


def _cubic(x1, f1, g1, x2, f2, g2, bounds):
    """lssr1_tr.py:280-309."""
    xmin, xmax = bounds
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)      # g1, g2 are 0-d tensors -> tensor math
    d2_sq = d1.pow(2) - g1 * g2
    if d2_sq >= 0:
        d2 = d2_sq.sqrt()
        if x1 <= x2:
            pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return torch.clamp(pos, xmin, xmax)
    return (xmin + xmax) / 2


class LSSR1TROracle:
    def __init__(self, *, lr=1.0, delta=1.0, min_delta=1e-3, max_delta=2.0, gamma=1e-3,
                 second_order=True, dogleg=False, mem_length=10, max_wolfe_iters=5,
                 max_zoom_iters=5, mu=0.9, tau_1=0.1, tau_2=0.25, tau_3=0.75, nu_1=0.25,
                 nu_2=0.5, nu_3=0.8, nu_4=1.2, tol=1e-8, norm_type=2, c_1=1e-4, c_2=0.9,
                 alpha_max=10.0, sync=False, paper_tr_update=True, dtype=torch.float32,
                 eig_sign="lapack", buf_dtype=None):
        if dogleg and not second_order:
            raise ValueError("Dogleg is only applicable in second-order mode")
        self.lr, self.delta, self.min_delta, self.max_delta = lr, delta, min_delta, max_delta
        self.second_order, self.dogleg = second_order, dogleg
        self.max_wolfe_iters, self.max_zoom_iters = max_wolfe_iters, max_zoom_iters
        self.mu, self.tau_1, self.tau_2, self.tau_3 = mu, tau_1, tau_2, tau_3
        self.nu_1, self.nu_2, self.nu_3, self.nu_4 = nu_1, nu_2, nu_3, nu_4
        self.tol = float(tol)
        self.c_1, self.c_2, self.alpha_max = c_1, c_2, alpha_max
        self.paper_tr_update = bool(paper_tr_update)
        self.hess = LSR1Memory(gamma=gamma, memory_length=mem_length, tol=self.tol, dtype=dtype)
        self.eig_sign = eig_sign
        # lssr1_tr.py:99-142, 205-229: without flat hooks the parameters and every evaluated
        # gradient pass through `flat_wk` / `flat_gk`, allocated in the dtype the first parameter had
        # AT CONSTRUCTION.  When the model is converted afterwards (the Trainer's .double(),
        # trainer.py:1203-1211) they are rounded to that dtype: pass it as ``buf_dtype``.
        self.buf_dtype = buf_dtype
        self.vk = None
        self.old_wk = None
        self.prev_grad = None
        self.n_evals = 0
        self.trace: list[dict] = []

    # lssr1_tr.py:247-278
    def _eval(self, fg, w, wk, p, alpha):
        w.copy_(wk + alpha * p)
        loss, grad = fg(w)
        if self.buf_dtype is not None:
            grad = grad.to(self.buf_dtype)                # _flatten_grads through flat_gk
        self.n_evals += 1
        return loss, grad, grad.dot(p.to(grad.dtype))

    # lssr1_tr.py:311-373
    def _zoom(self, fg, w, wk, p, a_lo, a_hi, f_lo, f_hi, d_lo, d_hi, f0, d0):
        g_last = None
        for i in range(self.max_zoom_iters):
            if i == 0:
                a_j = 0.5 * (a_lo + a_hi)
            else:
                lo = a_lo + 0.1 * (a_hi - a_lo)
                hi = a_hi - 0.1 * (a_hi - a_lo)
                a_j = _cubic(a_lo, f_lo, d_lo, a_hi, f_hi, d_hi, (lo, hi))
            f_j, g_j, d_j = self._eval(fg, w, wk, p, a_j)
            g_last = g_j
            if f_j > (f0 + self.c_1 * a_j * d0) or f_j >= f_lo:
                a_hi, f_hi, d_hi = a_j, f_j, d_j
            else:
                if abs(d_j) <= -self.c_2 * d0:
                    return a_j, f_j, g_j
                if d_j * (a_hi - a_lo) >= 0:
                    a_hi, f_hi, d_hi = a_lo, f_lo, d_lo
                a_lo, f_lo, d_lo = a_j, f_j, d_j
            if (a_hi - a_lo) * p.abs().max() < self.tol:
                break
        return a_lo, f_lo, g_last

    # lssr1_tr.py:375-460
    def _wolfe(self, fg, w, wk, p, f0, d0):
        a_prev, f_prev, d_prev = self.lr, f0, d0
        a_i = 0.5 * self.alpha_max
        g_prev = None
        for i in range(self.max_wolfe_iters):
            f_i, g_i, d_i = self._eval(fg, w, wk, p, a_i)
            if f_i > (f0 + self.c_1 * a_i * d0) or (i > 1 and f_i >= f_prev):
                return self._zoom(fg, w, wk, p, a_prev, a_i, f_prev, f_i, d_prev, d_i, f0, d0)
            if abs(d_i) <= -self.c_2 * d0:
                return a_i, f_i, g_i
            if d_i >= 0:
                return self._zoom(fg, w, wk, p, a_i, a_prev, f_i, f_prev, d_i, d_prev, f0, d0)
            a_prev, f_prev, d_prev, g_prev = a_i, f_i, d_i, g_i
            a_i = min(1.25 * a_i, self.alpha_max)
            if a_i >= self.alpha_max:
                break
        return a_prev, f_prev, g_prev

    # lssr1_tr.py:493-654
    def step(self, fg, w, loss=None, grad=None):
        if loss is None:
            loss, grad = fg(w)
        g = grad
        gn = torch.norm(g)
        rec = {"gn": float(gn), "delta_in": float(self.delta)}
        self.trace.append(rec)
        if gn <= self.tol:
            rec["converged"] = True
            return loss, g
        wk = w.clone() if self.buf_dtype is None else w.to(self.buf_dtype)   # _flatten_params through flat_wk
        updated = False
        if self.second_order and self.prev_grad is not None:          # :518-524
            sk, yk = wk - self.old_wk, g - self.prev_grad
            if sk.norm() > self.tol and yk.norm() > self.tol:
                rec["pair"] = self.hess.update(sk, yk)
                updated = True
        self.old_wk, self.prev_grad = wk.clone(), g.clone()
        k = len(self.hess)
        info = {}
        if self.second_order and k > 0:
            p_star, pred = solve_second_order(g, gn, self.delta, self.hess, self.tol, self.dogleg,
                                              eig_sign=self.eig_sign, info=info)
        else:
            p_star, pred = solve_first_order(g, gn, self.delta, self.tol)
        # dead momentum (:552-561): vk = mu*vk + (wk - old_wk) with old_wk == wk
        if self.vk is None:
            self.vk = torch.zeros_like(wk)                             # flat_vk: the buffer dtype
        vk = self.vk
        vk.mul_(self.mu).add_(wk - self.old_wk)
        vk_sq = vk.dot(vk)
        if vk_sq > self.tol:
            vk.mul_(min(1.0, self.delta / math.sqrt(float(vk_sq))))
        p = p_star + vk
        p_sq = p.dot(p)
        if p_sq > self.tol:                                            # :565-570
            p.mul_(min(1.0, self.delta / math.sqrt(float(p_sq))))
            p_sq = p.dot(p)
        self.vk = vk.clone()
        p = p.to(g.dtype)
        d0 = g.dot(p)
        if d0.item() > 0:                                              # :584-587
            p.mul_(-1)
            pred = pred * -1
            d0.mul_(-1)
        n0 = self.n_evals
        alpha, new_loss, new_g = self._wolfe(fg, w, wk, p, loss, d0)
        p_step = alpha * p
        w.copy_(wk + p_step)                                           # :605
        if abs(float(pred)) < self.tol:                                # :608-611
            rho = float("inf")
        else:
            rho = (loss - new_loss) / pred if (alpha > 0 and pred < 0) else 0.0
        s_sq = p_sq if alpha == 1.0 else p_step.dot(p_step)
        s_norm = math.sqrt(float(s_sq))
        if self.paper_tr_update:                                       # :621-639
            d_old = self.delta
            if rho < self.tau_2:
                d_new = min(self.nu_1 * d_old, self.nu_2 * s_norm)
            elif rho >= self.tau_3 and s_norm >= self.nu_3 * d_old:
                d_new = self.nu_4 * d_old
            else:
                d_new = d_old
            self.delta = max(self.min_delta, min(d_new, self.max_delta))
        else:                                                          # :641-652
            if rho > self.tau_2 and rho > self.tau_3 and s_norm >= 0.9 * self.delta:
                self.delta = min(self.max_delta, self.nu_4 * self.delta)
            else:
                self.delta = max(self.min_delta, self.nu_3 * self.delta)
        rec.update(pred=float(pred), rho=float(rho), alpha=float(alpha), k=k, d0=float(d0),
                   evals=self.n_evals - n0, delta_out=float(self.delta), p=p.clone(),
                   obs=info, s_norm=s_norm, pair_tried=updated)
        return new_loss, new_g.clone()
