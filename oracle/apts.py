"""Oracle: APTS outer iterations (data subdomains APTS_D / APTS_PINN weighting,
parameter subdomains APTS_P) with all ranks simulated in one process.
TEST INFRASTRUCTURE ONLY.

Reference: ``optimizers/apts_base.py:228-501``, ``apts_d.py:114-210``,
``apts_p.py:92-245``, ``apts_pinn.py:60-137``; hyper-parameter tables
``utility/optimizer_utils.py:27-151``.

Closures are per-rank functions on flat vectors: ``fg(w) -> (loss, grad)``.
The "all-reduce" is a sum over the simulated ranks in rank order (or a callable
passed as ``allreduce`` when driven by a real gloo group from the tests).
"""
from __future__ import annotations

import math

import torch

from .lssr1_tr import LSSR1TROracle
from .tr import TROracle


# ---------------------------------------------------------------- hparam tables
def apts_hparams(delta, min_delta, max_delta):
    """optimizer_utils.py:27-39."""
    return dict(delta=delta, min_delta=min_delta, max_delta=max_delta,
                nu_dec=0.25, nu_inc=0.75, inc_factor=1.2, dec_factor=0.9)


def tr_hparams(delta, min_delta, max_delta, norm_type=2, second_order=False, dogleg=False, tol=1e-6):
    """optimizer_utils.py:109-128 (mem_length hard-coded to 5)."""
    return dict(delta=delta, max_delta=max_delta, min_delta=min_delta, inc_factor=1.2,
                dec_factor=0.9, nu_dec=0.25, nu_inc=0.75, mem_length=5, norm_type=norm_type,
                second_order=second_order, dogleg=dogleg, tol=tol)


def loc_tr_hparams(delta, min_delta, max_delta, world_size, norm_type=2, second_order=False,
                   dogleg=False, tol=1e-6):
    """optimizer_utils.py:131-151: radii scaled by 1/world_size for the 2-norm."""
    sc = 1.0 / world_size if norm_type != math.inf else 1.0
    return dict(delta=delta * sc, max_delta=max_delta * sc, min_delta=min_delta * sc,
                inc_factor=1.5, dec_factor=0.6, nu_dec=0.3, nu_inc=0.7, mem_length=5,
                norm_type=norm_type, second_order=second_order, dogleg=dogleg, tol=tol)


def _lssr1_common(second_order, dogleg, mem_length, max_wolfe_iters, max_zoom_iters, tol,
                  norm_type, paper_tr_update):
    return dict(gamma=1e-3, second_order=second_order, dogleg=dogleg, mem_length=mem_length,
                max_wolfe_iters=max_wolfe_iters, max_zoom_iters=max_zoom_iters, mu=0.9,
                tau_1=0.1, tau_2=0.25, tau_3=0.75, nu_1=0.25, nu_2=0.5, nu_3=0.9, nu_4=1.2,
                tol=tol, norm_type=norm_type, c_1=1e-4, c_2=0.9, alpha_max=2.0,
                paper_tr_update=paper_tr_update)


def lssr1_hparams(delta, min_delta, max_delta, *, second_order=True, dogleg=False, mem_length=3,
                  max_wolfe_iters=5, max_zoom_iters=5, tol=1e-6, norm_type=2, paper_tr_update=False):
    """optimizer_utils.py:42-72."""
    d = _lssr1_common(second_order, dogleg, mem_length, max_wolfe_iters, max_zoom_iters, tol,
                      norm_type, paper_tr_update)
    d.update(delta=delta, min_delta=min_delta, max_delta=max_delta, sync=True)
    return d


def lssr1_loc_hparams(delta, min_delta, max_delta, world_size, *, second_order=True, dogleg=False,
                      mem_length=3, max_wolfe_iters=5, max_zoom_iters=5, tol=1e-6, norm_type=2,
                      paper_tr_update=False):
    """optimizer_utils.py:75-106."""
    sc = 1.0 / world_size if norm_type != math.inf else 1.0
    d = _lssr1_common(second_order, dogleg, mem_length, max_wolfe_iters, max_zoom_iters, tol,
                      norm_type, paper_tr_update)
    d.update(delta=delta * sc, min_delta=min_delta * sc, max_delta=max_delta * sc, sync=False)
    return d


def _make(kind, hp, **extra):
    if kind == "tr":
        return TROracle(**hp, **extra)
    if kind == "lssr1_tr":
        hp = {k: v for k, v in hp.items() if k != "sync"}
        extra = {k: v for k, v in extra.items() if k != "flat_hooks"}
        return LSSR1TROracle(**hp, **extra)
    raise ValueError(kind)


# ---------------------------------------------------------------- shared pieces
class _APTSBase:
    def __init__(self, w_init, n_ranks, glob_kind, glob_hp, loc_kind, loc_hp, *, delta, min_delta,
                 max_delta, nu_dec, nu_inc, inc_factor, dec_factor, glob_pass, norm_type,
                 max_loc_iters, max_glob_iters, tol, eig_sign):
        self.R = int(n_ranks)
        self.glob_pass, self.norm_type = bool(glob_pass), norm_type
        self.max_loc_iters, self.max_glob_iters = max_loc_iters, max_glob_iters
        self.tol = float(tol)
        self.min_delta, self.max_delta = float(min_delta), float(max_delta)
        self.nu_dec, self.nu_inc = float(nu_dec), float(nu_inc)
        self.inc_factor, self.dec_factor = float(inc_factor), float(dec_factor)
        self.glob_opt = _make(glob_kind, glob_hp, eig_sign=eig_sign)
        self.delta = self.glob_opt.delta                       # apts_base.py:142
        self.w = w_init.clone()                                # global model (identical on all ranks)
        self.grad_evals = 0.0
        self.trace: list[dict] = []

    # apts_base.py:294-318: average of the per-rank losses / gradients
    def _glob_fg(self, fgs):
        def fg(w):
            outs = [f(w) for f in fgs]
            loss = outs[0][0].clone()
            grad = outs[0][1].clone()
            for l, g in outs[1:]:
                loss = loss + l
                grad = grad + g
            if self.R > 1:
                loss = loss / self.R
                grad = grad / self.R
            self.grad_evals += 1.0
            return loss, grad
        return fg

    # apts_base.py:372-400
    def _step_loop(self, opt, fg, w, max_iters, loss, grad):
        for _ in range(max_iters):
            loss, grad = opt.step(fg, w, loss=loss, grad=grad)
            if grad.norm(p=self.norm_type) <= opt.tol:
                break
        return loss, grad

    # apts_base.py:432-501 (pred always supplied on the APTS_D / APTS_P paths)
    def _control(self, fg_glob, w0, f0, g0, step, pred):
        self.w.copy_(w0 + step)
        trial_loss, trial_grad = fg_glob(self.w)
        if abs(float(pred)) < self.tol:
            rho = float("inf")
        else:
            rho = (f0 - trial_loss) / pred
        rec = dict(pred=float(pred), rho=float(rho), step=step.clone())
        self.trace.append(rec)
        if pred < self.tol or rho < self.nu_dec:
            self.delta = max(self.delta * self.dec_factor, self.min_delta)
            self.w.copy_(w0)
            rec.update(accept=False, delta=self.delta)
            return f0, g0, self.delta
        if rho > self.nu_inc:
            self.delta = min(self.delta * self.inc_factor, self.max_delta)
        rec.update(accept=True, delta=self.delta)
        return trial_loss, trial_grad, self.delta


# ---------------------------------------------------------------- APTS_D
class APTSDOracle(_APTSBase):
    """apts_d.py:11-210 (+ apts_pinn.py:60-137 through ``weights``)."""

    def __init__(self, w_init, n_ranks, glob_kind, glob_hp, loc_kind, loc_hp, *, delta, min_delta,
                 max_delta, nu_dec, nu_inc, inc_factor, dec_factor, glob_pass=True, foc=True,
                 norm_type=2, max_loc_iters=3, max_glob_iters=3, tol=1e-6, eig_sign="lapack"):
        super().__init__(w_init, n_ranks, glob_kind, glob_hp, loc_kind, loc_hp, delta=delta,
                         min_delta=min_delta, max_delta=max_delta, nu_dec=nu_dec, nu_inc=nu_inc,
                         inc_factor=inc_factor, dec_factor=dec_factor, glob_pass=glob_pass,
                         norm_type=norm_type, max_loc_iters=max_loc_iters,
                         max_glob_iters=max_glob_iters, tol=tol, eig_sign=eig_sign)
        self.foc = bool(foc)
        self.loc_opts = [_make(loc_kind, loc_hp, eig_sign=eig_sign) for _ in range(self.R)]
        self.w_loc = [w_init.clone() for _ in range(self.R)]
        self.resid = [None] * self.R
        n = w_init.numel()
        if self.norm_type == math.inf:                         # apts_d.py:89-102
            sq = math.sqrt(n)
            self.glob_opt.min_delta = self.min_delta * sq
            self.glob_opt.max_delta = self.max_delta * sq
            self.glob_opt.delta = min(self.glob_opt.max_delta, self.delta * sq)
            for lo in self.loc_opts:
                lo.min_delta = self.min_delta * sq
                lo.max_delta = self.max_delta * sq
                lo.delta = min(lo.max_delta, self.delta * sq)

    def _loc_fg(self, r, fg_r):
        """apts_base.py:245-292; the FOC term changes the loss value only (Q7)."""
        def fg(w):
            loss, grad = fg_r(w)
            self._loc_evals[r] += 1
            if self.foc:
                diff = w - self.w
                if diff.norm() > self.tol:
                    res = self.resid[r]
                    if res is None:
                        res = torch.zeros_like(diff)
                    loss = loss + (res.to(diff.dtype) @ diff)
            return loss, grad
        return fg

    def step(self, fgs, weights=None, fgs_glob=None):
        """One outer iteration.  ``fgs[r]``: rank r's loss/grad on its local data;
        ``fgs_glob[r]`` (default ``fgs``): rank r's contribution to the global loss."""
        R = self.R
        self.grad_evals = 0.0
        self._loc_evals = [0.0] * R
        fg_glob = self._glob_fg(fgs if fgs_glob is None else fgs_glob)
        w0 = self.w.clone()
        f0, g0 = fg_glob(self.w)
        loc_fgs = [self._loc_fg(r, fgs[r]) for r in range(R)]
        lf0, lg0 = zip(*[loc_fgs[r](self.w_loc[r]) for r in range(R)])
        if self.foc:
            self.resid = [g0 - lg0[r] for r in range(R)]            # apts_d.py:186
        steps, reds = [], []
        for r in range(R):
            ll, _ = self._step_loop(self.loc_opts[r], loc_fgs[r], self.w_loc[r],
                                    self.max_loc_iters, lf0[r], lg0[r])
            steps.append(self.w_loc[r] - w0)
            reds.append(lf0[r] - ll)
        self.grad_evals += sum(self._loc_evals) / R                 # apts_base.py:413-417
        # additive Schwarz combine, apts_d.py:114-151
        if R > 1:
            step = steps[0].clone()
            wts = [1.0] * R if weights is None else weights
            pred = reds[0] * torch.tensor(float(wts[0]), dtype=step.dtype)
            for r in range(1, R):
                step = step + steps[r]
                pred = pred + reds[r] * torch.tensor(float(wts[r]), dtype=step.dtype)
            if self.norm_type == math.inf:
                step = step / R
        else:
            step, pred = steps[0], reds[0]
        loss, grad, self.glob_opt.delta = self._control(fg_glob, w0, f0, g0, step, pred)
        if self.glob_pass:
            loss, grad = self._step_loop(self.glob_opt, fg_glob, self.w, self.max_glob_iters, loss, grad)
        # apts_base.py:228-243
        self.delta = self.glob_opt.delta
        for r in range(R):
            self.loc_opts[r].delta = self.glob_opt.delta
            if self.norm_type == 2 and R > 1:
                self.loc_opts[r].delta /= R
            self.w_loc[r].copy_(self.w)
        return loss


# ---------------------------------------------------------------- APTS_P
class APTSPOracle(_APTSBase):
    """apts_p.py:11-245.  ``owned[r]`` = int64 flat indices (ascending, in model
    parameter order) of the elements rank r trains (apts_utils.py:83-124)."""

    def __init__(self, w_init, owned, glob_kind, glob_hp, loc_kind, loc_hp, *, delta, min_delta,
                 max_delta, nu_dec, nu_inc, inc_factor, dec_factor, glob_pass=True,
                 max_loc_iters=3, max_glob_iters=3, tol=1e-6, eig_sign="lapack", ctor_dtype=None):
        # ctor_dtype: dtype the model had when the reference optimizer was CONSTRUCTED, if it was
        # converted afterwards (trainer.py:1203-1211).  `_step_buffer` (apts_p.py:108) and, without flat
        # hooks, the global optimizer's flat buffers keep that dtype and round what passes through.
        self.ctor_dtype = ctor_dtype
        if ctor_dtype is not None and glob_kind == "lssr1_tr":
            glob_hp = dict(glob_hp, buf_dtype=ctor_dtype)
        super().__init__(w_init, len(owned), glob_kind, glob_hp, loc_kind, loc_hp, delta=delta,
                         min_delta=min_delta, max_delta=max_delta, nu_dec=nu_dec, nu_inc=nu_inc,
                         inc_factor=inc_factor, dec_factor=dec_factor, glob_pass=glob_pass,
                         norm_type=math.inf, max_loc_iters=max_loc_iters,
                         max_glob_iters=max_glob_iters, tol=tol, eig_sign=eig_sign)
        self.owned = owned
        self.loc_opts = [_make(loc_kind, loc_hp, eig_sign=eig_sign, flat_hooks=True) for _ in owned]
        self.w_loc = [w_init.clone() for _ in owned]
        self.sqrt_n = math.sqrt(w_init.numel())
        g = self.glob_opt                                           # apts_p.py:92-104
        g.max_delta = self.max_delta * self.sqrt_n
        g.min_delta = self.min_delta * self.sqrt_n
        g.delta = max(g.min_delta, min(g.max_delta, self.delta * self.sqrt_n))
        for lo in self.loc_opts:
            lo.max_delta, lo.min_delta, lo.delta = g.max_delta, g.min_delta, g.delta

    def step(self, fgs, fgs_glob=None):
        R = self.R
        self.grad_evals = 0.0
        evals = [0.0] * R
        fg_glob = self._glob_fg(fgs if fgs_glob is None else fgs_glob)
        w0 = self.w.clone()
        f0, g0 = fg_glob(self.w)
        lf0, reds = [], []
        for r in range(R):
            idx = self.owned[r]
            full = self.w_loc[r]

            def fg_own(w_own, _r=r, _idx=idx, _full=full):
                _full[_idx] = w_own
                loss, grad = fgs[_r](_full)
                evals[_r] += 1
                return loss, grad[_idx]

            w_own = full[idx].clone()
            l0, gr0 = fg_own(w_own)
            lf0.append(l0)
            ll, _ = self._step_loop(self.loc_opts[r], fg_own, w_own, self.max_loc_iters, l0, gr0)
            full[idx] = w_own
            reds.append(l0 - ll)
        self.grad_evals += sum(evals) / R
        merged = torch.zeros_like(w0)                               # apts_p.py:115-153
        for r in range(R):
            merged[self.owned[r]] = self.w_loc[r][self.owned[r]]
        step = merged - w0
        if self.ctor_dtype is not None:
            step = step.to(self.ctor_dtype)                         # torch.sub(..., out=self._step_buffer)
        step = step.clamp_(min=-self.delta, max=self.delta)         # :221-224
        pred = reds[0].clone()
        for r in range(1, R):
            pred = pred + reds[r]
        loss, grad, new_delta = self._control(fg_glob, w0, f0, g0, step, pred)
        self.delta = new_delta
        self.glob_opt.delta = min(self.glob_opt.max_delta, new_delta * self.sqrt_n)
        if self.glob_pass:
            loss, grad = self._step_loop(self.glob_opt, fg_glob, self.w, self.max_glob_iters, loss, grad)
        # apts_p.py:167-180
        self.delta = min(self.max_delta, max(self.min_delta, self.glob_opt.delta / self.sqrt_n))
        for r in range(R):
            self.loc_opts[r].delta = self.glob_opt.delta
            self.w_loc[r].copy_(self.w)
        return loss
