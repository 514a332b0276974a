"""L-S-SR1 trust-region optimizer with strong-Wolfe line search, drop-in for
``dd4ml.optimizers.lssr1_tr.LSSR1_TR`` (lssr1_tr.py:19-654).

Device work per step:

    lssr1_begin  s = w - old_w, y = g - prev_g, ALL reductions of update_memory and of the
                 sub-problem in one pass over (w, old_w, g, prev_g, S, Y); the candidate pair
                 goes to the spare ring slot; old_w <- w, prev_g <- g; the k-space kernel behind it runs the
                 acceptance chain, the Gram/gamma update, the L-SR1/OBS/dogleg solve, the
                 clip-to-radius and the descent flip
    tr_apply     p = cg g + sum_j (cS_j S_j + cY_j Y_j)            (direction, max|p|)
    per evaluation:  lssr1_trial  w = w_k + alpha p ;  <closure> ;  lssr1_eval  g.p, keep g

The Wolfe / zoom control flow is data dependent (lssr1_tr.py:311-460), so the host reads one
packed report per function evaluation (the reference: >= 10 syncs per step + 2 per evaluation).
Reference quirks kept: dead momentum (Q5), rho = 0 unless pred < 0 (Q6), never rejects.
"""
from __future__ import annotations

import math
import os
from typing import Callable

import numpy as np
import torch
import torch.distributed as dist
from torch.optim.optimizer import Optimizer

from .. import _lib
from .._device import DeviceState
from ..flat import get_flat
from ..utility.optimizer_utils import get_lssr1_tr_hparams
from .lsr1 import LSR1
from .tr import as_device_scalar


class LSSR1_TR(Optimizer):
    __name__ = "LSSR1_TR"

    @staticmethod
    def setup_LSSR1_TR_hparams(cfg):
        for k, v in get_lssr1_tr_hparams(cfg).items():
            setattr(cfg, k, v)
        return cfg

    def __init__(self, params, *, lr: float = 1.0, delta: float = 1.0, min_delta: float = 1e-3,
                 max_delta: float = 2.0, gamma: float = 1e-3, second_order: bool = True,
                 dogleg: bool = False, mem_length: int = 10, max_wolfe_iters: int = 5,
                 max_zoom_iters: int = 5, mu: float = 0.9, tau_1: float = 0.1, tau_2: float = 0.25,
                 tau_3: float = 0.75, nu_1: float = 0.25, nu_2: float = 0.5, nu_3: float = 0.8,
                 nu_4: float = 1.2, tol: float = 1e-8, norm_type: int = 2, c_1: float = 1e-4,
                 c_2: float = 0.9, alpha_max: float = 10.0, sync: bool = False,
                 paper_tr_update: bool = True, flat_grads_fn=None, flat_params_fn=None, flat_params=None):
        param_list = list(params)
        if not param_list:
            raise ValueError("Optimiser got an empty parameter list")
        if flat_params is not None:
            raise _lib.DD4MLError("dd4ml_b200 LSSR1_TR: the PMW `flat_params` path (APTS_IP) is out of scope")
        super().__init__(param_list, {"lr": lr})
        self.delta, self.min_delta, self.max_delta = delta, min_delta, max_delta
        self._gamma0 = gamma
        self.second_order, self.dogleg = second_order, dogleg
        if self.dogleg and not self.second_order:
            raise ValueError("Dogleg is only applicable in second-order mode")
        self.mem_length = mem_length
        self.max_wolfe_iters, self.max_zoom_iters = max_wolfe_iters, max_zoom_iters
        self.mu, self.tau_1, self.tau_2, self.tau_3 = mu, tau_1, tau_2, tau_3
        self.nu_1, self.nu_2, self.nu_3, self.nu_4 = nu_1, nu_2, nu_3, nu_4
        self.tol = float(tol)
        self.norm_type = 2                                        # lssr1_tr.py:94
        self.c_1, self.c_2, self.alpha_max = c_1, c_2, alpha_max
        self.sync = bool(sync)
        self.paper_tr_update = bool(paper_tr_update)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world_size = dist.get_world_size() if dist.is_initialized() else 1

        self.ps = [p for g in self.param_groups for p in g["params"]]
        self._fs = get_flat(self.ps)
        self._fs_version = self._fs.version
        device, dtype = self._fs.device, self._fs.dtype           # lssr1_tr.py:99-106
        # The reference allocates flat_wk / flat_gk in this dtype and routes the parameters and every
        # evaluated gradient through them (lssr1_tr.py:119-142, 205-229) unless flat hooks are given
        # (APTS_P's local optimizer).  If the model is converted afterwards (Trainer's .double(),
        # trainer.py:1203-1211) both keep being rounded to the construction dtype; reproduced below.
        self._ctor_dtype = dtype
        self._ref_hooks = False          # set by APTS_P for its local optimizer
        self.obs = None
        self.hess = LSR1(gamma=self._gamma0, memory_length=self.mem_length, device=device, dtype=dtype,
                         tol=self.tol)
        self._ds = self.hess.ds
        self._ds.set(second_order=1.0 if self.second_order else 0.0, dogleg=1.0 if self.dogleg else 0.0,
                     norm_inf=0.0, mem_len=self.mem_length, tol=self.tol)
        self._flat_grads_fn = flat_grads_fn
        self._flat_params_fn = flat_params_fn
        if flat_grads_fn is not None or flat_params_fn is not None:
            raise _lib.DD4MLError(
                "dd4ml_b200 LSSR1_TR: flat_grads_fn/flat_params_fn hooks are not needed -- pass the "
                "trainable parameters themselves (APTS_P lays them out contiguously)")
        # Folding the convergence read into the read of the first line-search evaluation (always at
        # alpha = alpha_max / 2) saves one host read per step, at the price of ONE closure call the
        # reference does not make when ||g|| <= tol (it returns before the line search, lssr1_tr.py:509).
        # Opt-in (DD4ML_B200_SPECULATIVE=1) for a stand-alone LSSR1_TR; the APTS classes switch it on for
        # their inner optimizers only when the model's forward pass has no side effects (no dropout RNG,
        # no BatchNorm running statistics): see APTS_Base._configure_speculation.
        self.speculative_first_eval = os.environ.get("DD4ML_B200_SPECULATIVE", "0") == "1"
        self._pre = None
        self._bufs = None
        self._has_prev = False
        self._slot = 0
        self._pushed_delta = None
        self.n_evals = 0
        self._elide_sync = False          # set by the APTS classes (see _sync_eval)
        self._last = None
        self.last_grad_norm = None
        self.last_grad_inf = None
        self._last_eval_norms = (0.0, 0.0)

    # lssr1_tr.py:597 keeps `self.gamma = self.hess.gamma` (a 0-d tensor) after every step; here it is
    # materialised on demand (building a device scalar from a host float is a synchronising copy)
    @property
    def gamma(self):
        return self.hess.gamma

    @property
    def last_report(self):
        """Diagnostics of the last step (built on demand, off the step's critical path)."""
        if self._last is None:
            return None
        rep, pred, rho, alpha, evals, d0 = self._last
        return dict(gn=rep.gn, pred=pred, rho=rho, alpha=alpha, k=int(rep.k), evals=evals, d0=d0,
                    case=int(rep.case_id), sigma=rep.sigma, pair_ok=rep.pair_ok, pair_tried=rep.pair_tried,
                    cyc_solve=rep.cyc_solve, cyc_eigh=rep.cyc_eigh, cyc_newton=rep.cyc_newton,
                    newton_iters=rep.newton_iters, jacobi_sweeps=rep.jacobi_sweeps)

    # ------------------------------------------------------------------ plumbing
    def zero_grad(self, set_to_none: bool = False) -> None:
        self._fs.zero_grad()

    def update_pytorch_lr(self) -> None:          # not in the reference class; harmless for APTS
        pass

    def _buffers(self, vt):
        if self._fs.version != self._fs_version or self._bufs is None or self._bufs["p"].dtype != vt:
            if self._bufs is not None and self._has_prev:
                # storage / dtype changed (e.g. model.double()): restart the pair history
                self._has_prev = False
            n, dev, wt = self._fs.n, self._fs.device, self._fs.dtype
            self._bufs = {
                "p": torch.zeros(n, dtype=vt, device=dev),
                # previous iterate / gradient, double buffered: lssr1_begin reads pair [cur] and writes pair
                # [1 - cur]; the host adopts the new pair unless the step turned out converged
                "old_w2": [torch.zeros(n, dtype=wt, device=dev) for _ in range(2)],
                "prev_g2": [torch.zeros(n, dtype=vt, device=dev) for _ in range(2)],
                "gs": [torch.zeros(n, dtype=vt, device=dev) for _ in range(3)],
            }
            self._cur = 0
            self._bufs["old_w"], self._bufs["prev_g"] = self._bufs["old_w2"][0], self._bufs["prev_g2"][0]
            self._fs_version = self._fs.version
        return self._bufs

    def _flatten_params(self) -> torch.Tensor:
        return self._fs.data.clone()

    def _flatten_grads(self) -> torch.Tensor:
        return self._fs.flat_grad().clone()

    # scalar arithmetic in the dtype the reference's 0-d tensors have
    def _f(self, x):
        return np.float32(x) if self._lt == torch.float32 else np.float64(x)

    # ------------------------------------------------------------------ lssr1_tr.py:175-203, 503-507, 272-276
    def _sync_eval(self, loss_t, grad):
        """`sync=True` on several ranks (LSSR1_TR as the top-level optimizer under DDP): the loss is
        averaged over the ranks (in fp64, as the reference does) and rank 0's gradient is broadcast,
        so every rank takes the same line-search decisions.  The reference also averages ||g|| and
        g.p; here both are reduced from the broadcast gradient on the device, which is the same
        number whenever the ranks hold the same gradient (DDP).  The APTS classes elide the whole
        exchange for their global optimizer: their closure has already all-reduced loss and
        gradient."""
        if not (self.sync and self.world_size > 1) or self._elide_sync:
            return loss_t
        buf = loss_t.detach().to(torch.float64)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        buf /= self.world_size
        dist.broadcast(grad, src=0)
        return buf.to(loss_t.dtype)

    def _emulate_ctor_dtype(self) -> bool:
        return (not self._ref_hooks) and self._fs.dtype != self._ctor_dtype

    # ------------------------------------------------------------------ evaluations
    def _evaluate(self, alpha, closure):
        """lssr1_tr.py:247-278: move to w_k + alpha p, evaluate, directional derivative."""
        lib, ds, b = self._ds.lib, self._ds, self._bufs
        n, st = self._fs.n, _lib.stream()
        vtc, wtc = _lib.dt(b["p"].dtype), _lib.dt(self._fs.dtype)
        alpha = float(alpha)
        if self._pre is not None and self._pre[0] == alpha:
            out, self._last_eval_norms, self._pre = self._pre[1], self._pre[2], None
            return out
        return self._evaluate_launch(alpha, closure, read=True)

    def _evaluate_launch(self, alpha, closure, read):
        lib, ds, b = self._ds.lib, self._ds, self._bufs
        n, st = self._fs.n, _lib.stream()
        vtc, wtc = _lib.dt(b["p"].dtype), _lib.dt(self._fs.dtype)
        if self._applied_alpha != alpha:
            lib.lssr1_trial(_lib.ptr(self._fs.data), _lib.ptr(b["old_w"]), _lib.ptr(b["p"]), alpha, n, vtc, wtc, st)
            self._applied_alpha = alpha
        loss = closure(compute_grad=True)
        self.n_evals += 1
        G = self._fs.flat_grad()
        if self._emulate_ctor_dtype():
            G = G.to(self._ctor_dtype)                 # _flatten_grads through the flat_gk buffer
        loss_t = self._sync_eval(as_device_scalar(loss, self._fs.device), G)
        if loss_t is not loss and self.sync and self.world_size > 1 and not self._elide_sync:
            loss = loss_t
        if G.dtype != b["p"].dtype:
            G = G.to(b["p"].dtype)
        self._slot = (self._slot + 1) % 3
        gs = b["gs"][self._slot]
        lib.lssr1_eval(ds.sp, ds.wp, _lib.ptr(G), _lib.ptr(b["p"]), _lib.ptr(gs), _lib.ptr(loss_t),
                       _lib.dt(loss_t.dtype), n, vtc, st)
        if not read:
            return loss, gs
        rep = ds.read()
        self._last_eval_norms = (rep.out_gg, rep.out_ginf)
        return (loss, self._f(rep.eval_loss)), gs, self._f(rep.deriv)

    @staticmethod
    def _cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
        """lssr1_tr.py:280-309 on host scalars."""
        if bounds is not None:
            xmin, xmax = bounds
        else:
            xmin, xmax = (x1, x2) if x1 <= x2 else (x2, x1)
        d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
        d2_sq = d1 * d1 - g1 * g2
        if d2_sq >= 0:
            d2 = np.sqrt(d2_sq)
            if x1 <= x2:
                pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
            else:
                pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
            return min(max(pos, xmin), xmax)
        return (xmin + xmax) / 2

    def _zoom(self, a_lo, a_hi, f_lo, f_hi, d_lo, d_hi, f0, d0, closure, pmax):
        """lssr1_tr.py:311-373.  f_* are (tensor, host value) pairs."""
        g_last = None
        c_1, c_2 = self.c_1, self.c_2
        for i in range(self.max_zoom_iters):
            if i == 0:
                a_j = 0.5 * (a_lo + a_hi)
            else:
                lo = a_lo + 0.1 * (a_hi - a_lo)
                hi = a_hi - 0.1 * (a_hi - a_lo)
                a_j = self._cubic_interpolate(a_lo, f_lo[1], d_lo, a_hi, f_hi[1], d_hi, bounds=(lo, hi))
            f_j, g_j, d_j = self._evaluate(a_j, closure)
            g_last = g_j
            if f_j[1] > (f0[1] + self._f(c_1 * a_j) * d0) or f_j[1] >= f_lo[1]:
                a_hi, f_hi, d_hi = a_j, f_j, d_j
            else:
                if abs(d_j) <= -self._f(c_2) * d0:
                    return a_j, f_j, g_j
                if d_j * (a_hi - a_lo) >= 0:
                    a_hi, f_hi, d_hi = a_lo, f_lo, d_lo
                a_lo, f_lo, d_lo = a_j, f_j, d_j
            if (a_hi - a_lo) * pmax < self.tol:
                break
        return a_lo, f_lo, g_last

    def _strong_wolfe_line_search(self, f0, d0, closure, pmax):
        """lssr1_tr.py:375-460."""
        c_1, c_2 = self.c_1, self.c_2
        a_prev, f_prev, d_prev, g_prev = self.defaults["lr"], f0, d0, None
        a_i = 0.5 * self.alpha_max
        for i in range(self.max_wolfe_iters):
            f_i, g_i, d_i = self._evaluate(a_i, closure)
            armijo = f_i[1] > (f0[1] + self._f(c_1 * a_i) * d0)
            if armijo or (i > 1 and f_i[1] >= f_prev[1]):
                return self._zoom(a_prev, a_i, f_prev, f_i, d_prev, d_i, f0, d0, closure, pmax)
            if abs(d_i) <= -self._f(c_2) * d0:
                return a_i, f_i, g_i
            if d_i >= 0:
                return self._zoom(a_i, a_prev, f_i, f_prev, d_i, d_prev, f0, d0, closure, pmax)
            a_prev, f_prev, d_prev, g_prev = a_i, f_i, d_i, g_i
            a_i = min(1.25 * a_i, self.alpha_max)
            if a_i >= self.alpha_max:
                break
        return a_prev, f_prev, g_prev

    # ------------------------------------------------------------------ the step
    def step(self, closure: Callable[[], torch.Tensor], **_):
        lib, ds = self._ds.lib, self._ds
        self._fs.ensure()                    # before reading gradients: .double() after construction
        dev = self._fs.device
        loss = _["loss"] if "loss" in _ else closure(compute_grad=True)
        g = _["grad"] if "grad" in _ else self._flatten_grads()
        g = g.detach()
        if not g.is_contiguous():
            g = g.contiguous()
        n = self._fs.n
        if g.numel() != n:
            raise _lib.DD4MLError(f"dd4ml_b200 LSSR1_TR: gradient has {g.numel()} elements, parameters {n}")
        b = self._buffers(g.dtype)
        h = self.hess
        h.ensure_storage(n)
        loss_t = self._sync_eval(as_device_scalar(loss, dev), g)
        if loss_t is not loss and self.sync and self.world_size > 1 and not self._elide_sync:
            loss = loss_t                         # the step returns the averaged value when nothing moves
        self._lt = loss_t.dtype
        # the radius updated on the host after the previous step travels with the begin launch
        rad = (-1.0, 0.0, 0.0)
        if self._pushed_delta != (self.delta, self.min_delta, self.max_delta):
            rad = (float(self.delta), float(self.min_delta), float(self.max_delta))
            self._pushed_delta = (self.delta, self.min_delta, self.max_delta)
        st = _lib.stream()
        vtc, wtc, htc = _lib.dt(g.dtype), _lib.dt(self._fs.dtype), _lib.dt(h.dtype)
        emul = self._emulate_ctor_dtype()
        w_in = self._fs.data
        if emul:                                     # wk = _flatten_params() through the flat_wk buffer
            w_in = b.get("w_round")
            if w_in is None:
                w_in = b["w_round"] = torch.empty_like(self._fs.data)
            w_in.copy_(self._fs.data.to(self._ctor_dtype))
        nxt = 1 - self._cur
        lib.lssr1_begin(ds.sp, ds.wp, _lib.ptr(w_in), _lib.ptr(b["old_w2"][self._cur]), _lib.ptr(b["old_w2"][nxt]),
                        _lib.ptr(g), _lib.ptr(b["prev_g2"][self._cur]), _lib.ptr(b["prev_g2"][nxt]),
                        1 if self._has_prev else 0, h.S_ptr, h.Y_ptr,
                        _lib.ptr(loss_t), _lib.dt(loss_t.dtype), n, h.ld, vtc, htc, wtc, self.mem_length, *rad, st)
        prev_pair = (b["old_w"], b["prev_g"])
        b["old_w"], b["prev_g"] = b["old_w2"][nxt], b["prev_g2"][nxt]      # w_k: the base of every trial point
        # The first trial point of the line search is w_k + 1.0 p (alpha_max / 2, lssr1_tr.py:397): when
        # it is evaluated speculatively, the direction kernel applies it on the way (w += p) instead
        # of a separate trial launch.  (begin has just stored old_w = w_k, so w + p == old_w + 1.0 p.)
        fuse_first = self.speculative_first_eval and 0.5 * self.alpha_max == 1.0 and not emul
        lib.tr_apply(ds.sp, ds.wp, _lib.ptr(g), h.S_ptr, h.Y_ptr, _lib.ptr(b["p"]),
                     _lib.ptr(self._fs.data) if fuse_first else None, n, h.ld, vtc, htc, wtc, self.mem_length, st)
        self._applied_alpha = 1.0 if fuse_first else 0.0
        self._pre = None
        pre = None
        if self.speculative_first_eval:
            pre = self._evaluate_launch(0.5 * self.alpha_max, closure, read=False)
        rep = ds.read()
        DeviceState.raise_for(rep.err)
        h.absorb(rep)
        if rep.converged:                                           # lssr1_tr.py:509
            self.last_grad_norm, self.last_grad_inf = rep.gn, rep.ginf
            if pre is not None:                                     # undo the speculative trial point
                lib.lssr1_trial(_lib.ptr(self._fs.data), _lib.ptr(b["old_w"]), _lib.ptr(b["p"]), 0.0, n, vtc, wtc, st)
            # the reference returns before touching old_wk / prev_grad (lssr1_tr.py:509-526): keep the
            # previous pair, the buffers lssr1_begin has just written are scratch again
            b["old_w"], b["prev_g"] = prev_pair
            return loss, g
        self._cur = nxt
        self._has_prev = True
        pred = self._f(rep.pred)
        d0 = self._f(rep.dphi0)
        psq, pmax = rep.psq, rep.pmax
        f0 = (loss, self._f(rep.loss))            # phi_0: copied into the report by lssr1_begin
        n0 = self.n_evals
        if pre is not None:
            n0 -= 1
            self._lt = loss_t.dtype
            self._pre = (0.5 * self.alpha_max, ((pre[0], self._f(rep.eval_loss)), pre[1], self._f(rep.deriv)),
                         (rep.out_gg, rep.out_ginf))
        alpha, new_f, new_g = self._strong_wolfe_line_search(f0, d0, closure, pmax)
        if self._applied_alpha != float(alpha):                       # lssr1_tr.py:605
            lib.lssr1_trial(_lib.ptr(self._fs.data), _lib.ptr(b["old_w"]), _lib.ptr(b["p"]), float(alpha), n,
                            vtc, wtc, st)
            self._applied_alpha = float(alpha)
        new_loss = new_f[0]
        # rho (lssr1_tr.py:608-611)
        if abs(float(pred)) < self.tol:
            rho = float("inf")
        else:
            rho = float((f0[1] - new_f[1]) / pred) if (alpha > 0 and pred < 0) else 0.0
        s_sq = psq if alpha == 1.0 else float(alpha) ** 2 * psq
        s_norm = math.sqrt(float(s_sq))
        if self.paper_tr_update:                                    # lssr1_tr.py:621-639
            d_old = self.delta
            if rho < self.tau_2:
                d_new = min(self.nu_1 * d_old, self.nu_2 * s_norm)
            elif rho >= self.tau_3 and s_norm >= self.nu_3 * d_old:
                d_new = self.nu_4 * d_old
            else:
                d_new = d_old
            self.delta = max(self.min_delta, min(d_new, self.max_delta))
        else:                                                       # lssr1_tr.py:641-652
            if rho > self.tau_2 and rho > self.tau_3 and s_norm >= 0.9 * self.delta:
                self.delta = min(self.max_delta, self.nu_4 * self.delta)
            else:
                self.delta = max(self.min_delta, self.nu_3 * self.delta)
        self._last = (rep, float(pred), rho, float(alpha), self.n_evals - n0, float(d0))
        # the returned gradient is always that of the last evaluation: its norms came with that
        # evaluation's report, so the caller's convergence test (apts_base.py:397) needs no reduction
        gg, ginf = self._last_eval_norms
        self.last_grad_norm = math.sqrt(gg) if new_g.dtype == torch.float64 else float(np.float32(math.sqrt(gg)))
        self.last_grad_inf = ginf
        return new_loss, new_g

    def state_dict(self):
        sd = super().state_dict()
        sd["dd4ml_b200"] = {"delta": self.delta, "hess": self.hess.state_dict(),
                            "has_prev": self._has_prev,
                            "old_w": None if self._bufs is None else self._bufs["old_w"].clone(),
                            "prev_g": None if self._bufs is None else self._bufs["prev_g"].clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        extra = state_dict.pop("dd4ml_b200", None)
        super().load_state_dict(state_dict)
        if extra is None:
            return
        self.hess.load_state_dict(extra["hess"])
        self.delta = float(extra["delta"])
        self._pushed_delta = None
        self._pre = None
        self._has_prev = bool(extra["has_prev"]) and extra["old_w"] is not None
        if extra["old_w"] is not None:
            b = self._buffers(extra["prev_g"].dtype)
            has_prev = self._has_prev          # _buffers() drops the history when it (re)allocates
            b["old_w"].copy_(extra["old_w"])
            b["prev_g"].copy_(extra["prev_g"])
            self._has_prev = has_prev
