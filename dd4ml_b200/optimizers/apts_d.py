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
        self._setup_p2p_combine()
        self._enable_device_control()
        from ..graphs import GraphedEval
        self._loc_eval = GraphedEval(self.loc_model, self.criterion, self._lfs)

    def _comm_dtype(self):
        return self._ctor_dtype

    def _setup_p2p_combine(self):
        """Symmetric-memory exchange buffers for the fused additive-Schwarz combine (csrc/k_apts.cu
        CombineOp): every rank's packed step lives in a buffer all ranks of the NVLink domain can load
        from, double-buffered by outer-iteration parity so one device-side barrier per iteration orders
        the peers' reads against the next write.  NCCL all-reduce remains the path when symmetric memory
        is unavailable (no NVLink P2P, non-NCCL backend, DD4ML_B200_P2P_COMBINE=0): the decision is taken
        collectively so that all ranks use the same exchange."""
        import ctypes
        import os
        self._p2p, self._p2p_note = None, "single subdomain"
        if not self._distributed:
            return
        ok = os.environ.get("DD4ML_B200_P2P_COMBINE", "1") != "0" and dist.get_backend() == "nccl"
        ws = dist.get_world_size()
        symm = None
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm      # noqa: F401
            except Exception as e:  # noqa: BLE001
                ok, self._p2p_note = False, f"symmetric memory unavailable: {e!r}"
        flag = torch.tensor([1 if (ok and ws <= 16) else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0:
            self._p2p_note = self._p2p_note if not ok else "disabled on another rank"
            if os.environ.get("DD4ML_B200_P2P_COMBINE", "1") == "0":
                self._p2p_note = "DD4ML_B200_P2P_COMBINE=0"
            return
        state = None
        try:
            n, dt = self._gfs.n, self._ctor_dtype
            L = (n + 2 + 3) // 4 * 4
            t = symm.empty(2 * L, dtype=dt, device=self.device)
            t.zero_()
            hdl = symm.rendezvous(t, dist.group.WORLD)
            es = t.element_size()
            tables = [(ctypes.c_void_p * ws)(*[int(p) + b * L * es for p in hdl.buffer_ptrs]) for b in (0, 1)]
            state = dict(t=t, hdl=hdl, L=L, tables=tables, bufs=[t[:n + 2], t[L:L + n + 2]], it=0, ws=ws,
                         tail=torch.zeros(2, dtype=dt, device=self.device))
        except Exception as e:  # noqa: BLE001
            self._p2p_note = f"rendezvous failed: {e!r}"
        flag = torch.tensor([1 if state is not None else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 1:
            self._p2p, self._p2p_note = state, "peer-memory combine over symmetric memory"

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

        p2p = self._p2p
        if p2p is not None:
            par = p2p["it"] & 1
            p2p["it"] += 1
            comm = p2p["bufs"][par]
        else:
            comm = b["comm"]
        lib.apts_pack(_lib.ptr(comm), _lib.ptr(self._lfs.data), _lib.ptr(b["w0"]), _lib.ptr(self.init_loc_loss),
                      _lib.ptr(loc_loss), _lib.dt(self.init_loc_loss.dtype),
                      float(weight) if self.nr_models > 1 else 1.0, float(self.loc_grad_evals), n, wt,
                      _lib.dt(comm.dtype), st)
        scale = 1.0
        if self._distributed and self.norm_type == math.inf:
            scale = 1.0 / self.nr_models
        use(glob_io, on_glob)
        es = comm.element_size()
        if p2p is not None:
            # additive Schwarz (apts_d.py:144) + w = w0 + step (apts_base.py:445) in ONE kernel over NVLink
            # peer memory; the barrier orders it behind every rank's pack kernel
            self._p2p_exchange(comm, p2p, par, scale, b, n, wt, st)
            tail = p2p["tail"]
            loss, grad, delta, evals_sum, gn = self._control(
                None, 0, scale, 0.0, self.init_glob_loss, tail.data_ptr(), _lib.dt(tail.dtype),
                tail.data_ptr() + es, pre_applied=True)
        else:
            if self._distributed:
                apts_base.all_reduce(comm, dist.ReduceOp.SUM)         # additive Schwarz (apts_d.py:144)
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

    def _p2p_exchange(self, comm, p2p, par, scale, b, n, wt, st):
        lib = self._ctl.lib
        if apts_base.COMM_PROF is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        p2p["hdl"].barrier(channel=0)
        lib.apts_combine(_lib.ptr(self._gfs.data), _lib.ptr(b["w0"]), p2p["tables"][par], p2p["ws"], float(scale),
                         _lib.ptr(p2p["tail"]), n, wt, _lib.dt(comm.dtype), st)
        if apts_base.COMM_PROF is not None:
            e1.record()
            apts_base.COMM_PROF.append((e0, e1, -(n + 2)))      # negative payload marks the fused combine

    def step(self, inputs, labels, inputs_d=None, labels_d=None, hNk=None):
        self.inputs_d, self.labels_d, self.hNk = inputs_d, labels_d, hNk
        return self._outer((inputs, labels), (inputs, labels))
