"""Oracle: plain trust-region optimizer step (reference ``optimizers/tr.py:138-222``).
TEST INFRASTRUCTURE ONLY.

Works on a flat parameter vector ``w`` (modified in place) and a closure
``fg(w) -> (loss 0-d tensor, flat grad)``.
"""
from __future__ import annotations

import torch

from .lsr1 import LSR1Memory
from .subproblem import solve_first_order, solve_second_order


class TROracle:
    def __init__(self, *, delta=0.1, min_delta, max_delta, nu_dec, nu_inc, inc_factor,
                 dec_factor, norm_type=2, tol=1e-6, second_order=False, dogleg=False,
                 mem_length=10, flat_hooks=False, eig_sign="lapack", **_ignored):
        if dogleg and not second_order:
            raise ValueError("Dogleg is only applicable in second-order mode")  # tr.py:40
        self.delta = delta
        self.min_delta, self.max_delta = min_delta, max_delta
        self.nu_dec, self.nu_inc = nu_dec, nu_inc
        self.inc_factor, self.dec_factor = inc_factor, dec_factor
        self.norm_type, self.tol = norm_type, float(tol)
        self.second_order, self.dogleg = bool(second_order), bool(dogleg)
        # tr.py:69,76-78: the gradient buffer and the LSR1 memory are fp32 unless the
        # flat_grads_fn hook supplies the gradient (APTS_P), SURVEY Q4.
        self.flat_hooks = flat_hooks
        self.hess = LSR1Memory(gamma=1.0, memory_length=mem_length, tol=self.tol) if second_order else None
        self.eig_sign = eig_sign
        self.trace: list[dict] = []

    def _grad(self, g):                                   # tr.py:88-104
        return g.clone() if self.flat_hooks else g.to(torch.float32).clone()

    def step(self, fg, w, loss=None, grad=None):
        if loss is None:
            loss, g0 = fg(w)
            grad = self._grad(g0)
        gn = torch.norm(grad, p=self.norm_type)
        rec = {"gn": float(gn), "delta_in": float(self.delta)}
        self.trace.append(rec)
        if gn <= self.tol:                                 # tr.py:147
            rec["converged"] = True
            return loss, grad
        k = len(self.hess) if self.second_order else 0
        info = {}
        if self.second_order and k > 0:                    # tr.py:160
            step, pred = solve_second_order(grad, gn, self.delta, self.hess, self.tol,
                                            self.dogleg, eig_sign=self.eig_sign, info=info)
        else:
            step, pred = solve_first_order(grad, gn, self.delta, self.tol)
        step_norm = step.norm()
        w.add_(step)                                       # tr.py:179 (_apply_update)
        trial_loss, tg = fg(w)
        trial_grad = self._grad(tg)
        if abs(float(pred)) < self.tol:                    # tr.py:186
            rho = float("inf")
        else:
            rho = (loss - trial_loss) / pred
        rec.update(pred=float(pred), rho=float(rho), step_norm=float(step_norm), k=k,
                   step=step.clone(), obs=info)
        if rho > self.nu_dec:                              # accept, tr.py:191
            rec["accept"] = True
            if self.second_order:
                yk = trial_grad - grad
                if step_norm > self.tol and yk.norm() > self.tol:
                    rec["pair"] = self.hess.update(step.clone(), yk.clone())
            if rho > self.nu_inc and step_norm >= 0.9 * self.delta:
                self.delta = min(self.max_delta, self.inc_factor * self.delta)
            rec["delta_out"] = float(self.delta)
            return trial_loss, trial_grad
        rec["accept"] = False
        w.add_(step, alpha=-1.0)                           # tr.py:216
        self.delta = max(self.min_delta, self.dec_factor * self.delta)
        rec["delta_out"] = float(self.delta)
        return loss, grad
