"""APTS with data subdomains, drop-in for ``dd4ml.optimizers.apts_d.APTS_D`` (apts_d.py:4-210).

One outer iteration on each rank (= one B200 = one data subdomain):

    w0 <- w                                   save the initial global parameters
    global closure, local closure             [g || loss] all-reduce inside the global one
    apts_residual                             g0 <- G_glob, gl0 <- G_loc, resid <- G_glob - G_loc
    <= max_loc_iters local TR / LSSR1_TR steps on the local clone (no communication)
    apts_pack                                 comm <- [w_loc - w0 || (f_loc0 - f_loc) * weight || evals]
    all-reduce(SUM) of n+2 elements           the additive-Schwarz combination (apts_d.py:144)
    apts_trial, global closure, apts_control  trust-region control of the combined step
    <= max_glob_iters global steps, then w_loc <- w and the radius bookkeeping
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from .. import _lib
from ..flat import get_flat
from ..utility.apts_utils import clone_model
from . import apts_base
from .apts_base import APTS_Base
from .tr import as_device_scalar


class APTS_D(APTS_Base):
    __name__ = "APTS_D"

    @staticmethod
    def setup_APTS_hparams(config):
        return APTS_Base.setup_APTS_hparams(config)

    def __init__(self, params, model=None, delta=None, min_delta=None, max_delta=None, nu_dec=None,
                 nu_inc=None, inc_factor=None, dec_factor=None, criterion=None, device=None,
                 nr_models=None, glob_opt=None, glob_opt_hparams=None, loc_opt=None, loc_opt_hparams=None,
                 *, glob_pass=True, foc=True, soc=False, norm_type=2, max_loc_iters=3, max_glob_iters=3,
                 tol=1e-6):
        super().__init__(params, model=model, delta=delta, min_delta=min_delta, max_delta=max_delta,
                         nu_dec=nu_dec, nu_inc=nu_inc, inc_factor=inc_factor, dec_factor=dec_factor,
                         criterion=criterion, device=device, nr_models=nr_models, glob_opt=glob_opt,
                         glob_opt_hparams=glob_opt_hparams, loc_opt=loc_opt, loc_opt_hparams=loc_opt_hparams,
                         glob_pass=glob_pass, norm_type=norm_type, max_loc_iters=max_loc_iters,
                         max_glob_iters=max_glob_iters, tol=tol)
        self.foc, self.soc = bool(foc), bool(soc)
        self.loc_model = clone_model(model)
        self._lfs = get_flat(list(self.loc_model.parameters()))
        self.loc_opt = loc_opt(self.loc_model.parameters(), **loc_opt_hparams)
        self.loc_closure = self.foc_loc_closure if self.foc else self.non_foc_loc_closure
        self.n_global, self.n_local = self._gfs.n, self._lfs.n
        if self.norm_type == math.inf:                               # apts_d.py:89-102
            self.sqrt_n_global, self.sqrt_n_local = math.sqrt(self.n_global), math.sqrt(self.n_local)
            self.glob_opt.min_delta = self.min_delta * self.sqrt_n_global
            self.glob_opt.max_delta = self.max_delta * self.sqrt_n_global
            self.loc_opt.min_delta = self.min_delta * self.sqrt_n_local
            self.loc_opt.max_delta = self.max_delta * self.sqrt_n_local
            self.glob_opt.delta = min(self.glob_opt.max_delta, self.delta * self.sqrt_n_global)
            self.loc_opt.delta = min(self.loc_opt.max_delta, self.delta * self.sqrt_n_local)
        self._has_buffers = any(True for _ in self.loc_model.buffers())
        self._enable_device_control()
        from ..graphs import GraphedEval
        self._loc_eval = GraphedEval(self.loc_model, self.criterion, self._lfs)

    def _comm_dtype(self):
        return self._ctor_dtype

    # ------------------------------------------------------------------ apts_base.py:228-243
    @torch.no_grad()
    def sync_glob_to_loc(self):
        if self._device_control:        # radii handed over on the device, read once at the end of step()
            scale = 1.0 / self.nr_models if (self.norm_type == 2 and self.nr_models > 1) else 1.0
            self._link(self._ctl, self.glob_opt._ds)
            self._link(self.loc_opt._ds, self.glob_opt._ds, scale)
            self.loc_opt._delta_stale = True
        else:
            self.delta = self.glob_opt.delta
            self.update_pytorch_lr()
            self.loc_opt.delta = self.glob_opt.delta
            if self.norm_type == 2 and self.nr_models > 1:
                self.loc_opt.delta /= self.nr_models
            if hasattr(self.loc_opt, "update_pytorch_lr"):
                self.loc_opt.update_pytorch_lr()
        # load_state_dict(global) as ONE flat device copy (+ module buffers such as BatchNorm stats)
        if self._gfs.ensure() | self._lfs.ensure():
            self._buffers()
        self._lfs.data.copy_(self._gfs.data)
        if self._has_buffers:
            base = self.model.module if hasattr(self.model, "module") else self.model
            for lb, gb in zip(self.loc_model.buffers(), base.buffers()):
                lb.copy_(gb)

    @torch.no_grad()
    def aggregate_loc_steps_and_losses(self, step, loc_red, *, weight: float = 1.0):
        """apts_d.py:114-151 (public form): sum of the local steps and of the weighted local
        reductions over the subdomains with one coalesced all-reduce."""
        if self.nr_models > 1 and self._distributed:
            comm = self._buffers()["comm"]
            n = step.numel()
            comm[:n].copy_(step.view(-1))
            comm[n] = as_device_scalar(loc_red, self.device).to(comm.dtype) * weight
            dist.all_reduce(comm[: n + 1], op=dist.ReduceOp.SUM)
            step.copy_(comm[:n].view_as(step))
            loc_red = comm[n].clone().unsqueeze(0)
            if self.norm_type == math.inf:
                step /= self.nr_models
        return step, loc_red

    # ------------------------------------------------------------------ apts_d.py:153-210
    def _outer(self, glob_io, loc_io, weight=1.0, on_glob=None, on_loc=None):
        """Shared by APTS_D.step and APTS_PINN.step: ``glob_io`` / ``loc_io`` are the
        (inputs, labels) pairs of the global and the local closure."""
        lib = self._ctl.lib
        self.grad_evals, self.loc_grad_evals = 0.0, 0.0
        if self._resync:                 # first step after load_state_dict
            self.sync_glob_to_loc()
            self._resync = False
        if self._gfs.ensure() | self._lfs.ensure():
            pass
        b = self._buffers()
        n, wt, st = self._gfs.n, _lib.dt(self._gfs.dtype), _lib.stream()
        # gradient-norm caches refer to buffers this iteration overwrites in place
        self.glob_opt._known_gn = self.loc_opt._known_gn = None
        b["w0"].copy_(self._gfs.data)                                # init_glob_flat
        self.init_glob_flat = b["w0"]

        def use(io, hook):
            self.inputs, self.labels = io
            if hook is not None:
                hook()

        use(glob_io, on_glob)
        self.init_glob_loss = self.glob_closure_main(compute_grad=True)
        Gg = self._gfs.flat_grad()
        use(loc_io, on_loc)
        self.init_loc_loss = as_device_scalar(self.loc_closure(compute_grad=True), self.device)
        Gl = self._lfs.flat_grad()
        lib.apts_residual(_lib.ptr(Gg), _lib.ptr(Gl), _lib.ptr(b["g0"]), _lib.ptr(b["gl0"]),
                          _lib.ptr(b["resid"]) if self.foc else None, n, wt, st)
        self.init_glob_grad, self.init_loc_grad = b["g0"], b["gl0"]
        self.resid = b["resid"]

        loc_loss, _ = self.loc_steps(self.init_loc_loss, self.init_loc_grad)
        loc_loss = as_device_scalar(loc_loss, self.device, self.init_loc_loss.dtype)

        comm = b["comm"]
        lib.apts_pack(_lib.ptr(comm), _lib.ptr(self._lfs.data), _lib.ptr(b["w0"]), _lib.ptr(self.init_loc_loss),
                      _lib.ptr(loc_loss), _lib.dt(self.init_loc_loss.dtype),
                      float(weight) if self.nr_models > 1 else 1.0, float(self.loc_grad_evals), n, wt,
                      _lib.dt(comm.dtype), st)
        scale = 1.0
        if self._distributed:
            apts_base.all_reduce(comm, dist.ReduceOp.SUM)             # additive Schwarz (apts_d.py:144)
            if self.norm_type == math.inf:
                scale = 1.0 / self.nr_models
        use(glob_io, on_glob)
        es = comm.element_size()
        loss, grad, delta, evals_sum, gn = self._control(
            comm, 0, scale, 0.0, self.init_glob_loss, comm.data_ptr() + n * es, _lib.dt(comm.dtype),
            comm.data_ptr() + (n + 1) * es)
        if self._device_control:
            self._link(self.glob_opt._ds, self._ctl)                  # glob_opt.delta <- control delta
            self.glob_opt._delta_stale = True
            if self.glob_pass:
                loss, grad = self.glob_steps(loss, grad)
            self.sync_glob_to_loc()
            delta, evals_sum, gn = self._read_control()               # the only host read of the iteration
        else:
            self.glob_opt.delta = delta
            if self.glob_pass:
                self.glob_opt._known_gn = None if gn is None else (id(grad), grad._version, grad.data_ptr(), gn)
                loss, grad = self.glob_steps(loss, grad)
            self.sync_glob_to_loc()
        self.grad_evals += evals_sum / (self.nr_models if self._distributed else 1)   # apts_base.py:413-417
        return loss

    def step(self, inputs, labels, inputs_d=None, labels_d=None, hNk=None):
        self.inputs_d, self.labels_d, self.hNk = inputs_d, labels_d, hNk
        return self._outer((inputs, labels), (inputs, labels))
