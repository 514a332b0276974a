"""Shared machinery of the APTS optimizers, drop-in for ``dd4ml.optimizers.apts_base``
(apts_base.py:36-501).

Device-side design:
  * global and local model parameters / gradients are views of flat buffers (flat.py);
  * the global gradient and loss are averaged over the data subdomains with ONE in-place
    NCCL all-reduce of [g || loss] on the flat gradient buffer (the reference: flatten,
    all-reduce, scatter back + a second scalar all-reduce, apts_base.py:306-317);
  * the first-order-consistency term is one fused reduction (apts_base.py:262-292);
  * ``control_step`` (apts_base.py:432-501) is two kernels around the trial evaluation:
    ``apts_trial`` (w = w0 + step) and ``apts_control`` (rho test on the device; accepted:
    g0 <- trial gradient, rejected: w <- w0; radius update), then one report read.
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist
from torch.optim.optimizer import Optimizer

from .. import _lib
from .._device import DeviceState
from ..flat import get_flat
from ..graphs import GraphedEval
from ..utility.optimizer_utils import (get_apts_hparams, get_loc_tr_hparams, get_loc_tradam_hparams,
                                       get_lssr1_loc_tr_hparams, get_lssr1_tr_hparams, get_tr_hparams)
from .lssr1_tr import LSSR1_TR
from .tr import TR, as_device_scalar
from .tradam import TRAdam


COMM_PROF = None      # bench.py: list of (start_event, end_event, numel) per collective when profiling


def all_reduce(t, op):
    """Every collective of the APTS path goes through here (NCCL on the current stream)."""
    if COMM_PROF is None:
        dist.all_reduce(t, op=op)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.all_reduce(t, op=op)
    e1.record()
    COMM_PROF.append((e0, e1, t.numel()))


class APTS_Base(Optimizer):
    __name__ = "APTS_Base"

    @staticmethod
    def setup_APTS_hparams(config):
        """apts_base.py:40-83: resolve the optimizer names to the B200 classes + hparam tables.  `tradam` and
        `sgd` are accepted as local optimizers exactly as the reference accepts them (apts_base.py:50-55): under
        APTS the reference drives TRAdam with closures that do not back-propagate (:379-382) so its local phase
        leaves the parameters unchanged, and a `torch.optim.SGD` local optimizer raises TypeError on the first
        local step (:384-390) -- both behaviours are kept (tests/test_gpu_optimizers.py)."""
        glob_map = {"tr": (TR, get_tr_hparams), "lssr1_tr": (LSSR1_TR, get_lssr1_tr_hparams)}
        loc_map = {"tr": (TR, get_loc_tr_hparams), "lssr1_tr": (LSSR1_TR, get_lssr1_loc_tr_hparams),
                   "tradam": (TRAdam, get_loc_tradam_hparams), "sgd": (torch.optim.SGD, lambda _: {"lr": 0.01})}
        if not isinstance(config.glob_opt, str):
            raise ValueError("glob_opt must be a string specifying the optimizer type, e.g., 'tr' or 'lssr1_tr'.")
        try:
            config.glob_opt, fn = glob_map[config.glob_opt.lower()]
        except KeyError:
            raise ValueError(f"Unknown glob_opt: {config.glob_opt}")
        config.glob_opt_hparams = fn(config)
        loc_value = getattr(config, "loc_opt", None)
        if not isinstance(loc_value, str):
            raise ValueError("loc_opt must be a string specifying the optimizer type, e.g., 'tr' or 'lssr1_tr'.")
        try:
            config.loc_opt, fn = loc_map[loc_value.lower()]
        except KeyError:
            raise ValueError(f"Unknown loc_opt: {loc_value}")
        config.loc_opt_hparams = fn(config)
        config.apts_params = get_apts_hparams(config)
        return config

    def __init__(self, params, model=None, delta=None, min_delta=None, max_delta=None, nu_dec=None,
                 nu_inc=None, inc_factor=None, dec_factor=None, criterion=None, device=None,
                 nr_models=None, glob_opt=None, glob_opt_hparams=None, loc_opt=None, loc_opt_hparams=None,
                 *, glob_pass=True, norm_type=None, max_loc_iters=None, max_glob_iters=None, tol=1e-6):
        super().__init__(params, {"lr": delta})
        self.glob_pass = bool(glob_pass)
        self.norm_type = norm_type
        self.max_loc_iters = max_loc_iters
        self.max_glob_iters = max_glob_iters
        self.tol = float(tol)
        self.delta = float(delta)
        self.min_delta = float(min_delta) if min_delta is not None else 1e-3
        self.nu_dec = float(nu_dec) if nu_dec is not None else 0.25
        self.nu_inc = float(nu_inc) if nu_inc is not None else 0.75
        self.inc_factor = float(inc_factor) if inc_factor is not None else 1.2
        self.dec_factor = float(dec_factor) if dec_factor is not None else 0.9
        self.max_delta = float(max_delta) if max_delta is not None else 2.0
        self.nr_models = nr_models
        self.criterion = criterion
        self.dataset_len = 0
        self.model = model
        self._gfs = get_flat(list(model.parameters()))       # bind BEFORE the global optimizer so it reuses it
        self.device = self._gfs.device
        # dtype of the model at construction: the reference pre-allocates `_step_buffer` (and sizes its
        # coalesced all-reduce buffer after it) here (apts_d.py:112, apts_p.py:108), so when the Trainer
        # converts the model to float64 afterwards the step keeps being rounded to this dtype
        self._ctor_dtype = self._gfs.dtype
        self.glob_opt = glob_opt(self.model.parameters(), **glob_opt_hparams)
        if isinstance(self.glob_opt, LSSR1_TR):
            self.glob_opt._elide_sync = True    # glob_closure_main all-reduces loss and gradient already
        self.loc_model = None
        self.loc_opt = None
        self.loc_closure = None
        self.loc_closure_d = None
        self.delta = self.glob_opt.delta                    # apts_base.py:142
        self.grad_evals, self.loc_grad_evals = 0.0, 0.0
        self._ctl = DeviceState(self.device)
        self._ctl_pushed = {}
        self._distributed = bool(nr_models and nr_models > 1 and dist.is_available() and dist.is_initialized())
        self._is_ddp = isinstance(model, torch.nn.parallel.DistributedDataParallel)
        self._bufs = None
        self._bufs_version = None
        self.last_control = None
        # CUDA-graph replay of the (untouched) PyTorch forward/backward of the closures
        self._glob_eval = None if self._is_ddp else GraphedEval(self.model, self.criterion, self._gfs)
        self._loc_eval = None          # created by the subclass once the local model exists
        self._device_control = False   # see _enable_device_control
        self._resync = False           # set by load_state_dict

    # ------------------------------------------------------------------ checkpointing
    def state_dict(self):
        """Radius, evaluation counters and the inner optimizers' state (radius, L-SR1 memory).  The
        local model is re-synchronised from the global one at the start of every outer iteration,
        so it is not part of the state."""
        sd = super().state_dict()
        sd["dd4ml_b200"] = {"delta": float(self.delta), "grad_evals": float(self.grad_evals),
                            "loc_grad_evals": float(self.loc_grad_evals),
                            "glob": self.glob_opt.state_dict(),
                            "loc": None if self.loc_opt is None else self.loc_opt.state_dict()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        extra = state_dict.pop("dd4ml_b200", None)
        super().load_state_dict(state_dict)
        if extra is None:
            return
        self.glob_opt.load_state_dict(extra["glob"])
        if extra["loc"] is not None and self.loc_opt is not None:
            self.loc_opt.load_state_dict(extra["loc"])
        self.delta = float(extra["delta"])
        self.grad_evals = float(extra["grad_evals"])
        self.loc_grad_evals = float(extra["loc_grad_evals"])
        self._ctl_pushed = {}
        # the local clone mirrors the global model between outer iterations (sync_glob_to_loc at the
        # end of step()); the caller may restore the model before or after this call, so the next
        # step() re-establishes the invariant first
        self._resync = True

    def _enable_device_control(self):
        """TR/TR configurations run an outer iteration with ONE host read at its end: the inner TR
        steps, the control step and the radius hand-over between the APTS / global / local state
        blocks all stay on the device (DD4ML_B200_DEVICE_CONTROL=0 restores one read per step)."""
        env_on = os.environ.get("DD4ML_B200_DEVICE_CONTROL", "1") != "0"
        loc_tr = type(self.loc_opt) is TR and self.loc_opt._flat_params_fn is None
        on = env_on and type(self.glob_opt) is TR and loc_tr and self.glob_opt._flat_params_fn is None
        self._device_control = on
        self.glob_opt.device_control = on if type(self.glob_opt) is TR else False
        # a local TR never needs the host inside its loop, whatever the global optimizer is
        # (its radius is re-derived from the global one at the end of every outer iteration)
        if type(self.loc_opt) is TR:
            self.loc_opt.device_control = env_on and loc_tr
        self._configure_speculation()

    @staticmethod
    def closure_has_side_effects(model) -> bool:
        """True when an extra forward pass would be observable: dropout consumes RNG state, BatchNorm
        updates running statistics.  Models that draw random numbers without using
        nn.Dropout modules can declare it with an attribute `stochastic_forward = True`."""
        base = model.module if hasattr(model, "module") else model
        if getattr(base, "stochastic_forward", False):
            return True
        for m in base.modules():
            if isinstance(m, torch.nn.modules.dropout._DropoutNd) and m.p > 0:
                return True
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.track_running_stats:
                return True
            if isinstance(m, (torch.nn.RNNBase, torch.nn.MultiheadAttention)) and getattr(m, "dropout", 0) > 0:
                return True
        return False

    def _configure_speculation(self):
        """LSSR1_TR inner optimizers may launch the first line-search evaluation before the convergence
        test has been read (one host read less per step).  The only difference from the reference's
        closure-call sequence is one extra evaluation when ||g|| <= tol, so it is enabled only for
        models whose forward pass has no side effects, and never when DD4ML_B200_SPECULATIVE=0 or
        DD4ML_B200_DEVICE_CONTROL=0."""
        on = (os.environ.get("DD4ML_B200_DEVICE_CONTROL", "1") != "0"
              and os.environ.get("DD4ML_B200_SPECULATIVE", "1") != "0"
              and not self.closure_has_side_effects(self.model))
        for o in (self.glob_opt, self.loc_opt):
            if isinstance(o, LSSR1_TR):
                o.speculative_first_eval = on
        return on

    def _link(self, dst, src, scale=1.0, lo=-math.inf, hi=math.inf):
        """dst.delta <- clamp(src.delta * scale, lo, hi) on the device."""
        off = self._ctl.lay.off("delta")
        self._ctl.lib.state_link(dst.sp, off, src.sp, off, float(scale), float(lo), float(hi), _lib.stream())

    # ------------------------------------------------------------------ buffers
    def _buffers(self):
        ver = (self._gfs.version, self._gfs.dtype)
        if self._bufs is None or self._bufs_version != ver:
            n, dev, wt = self._gfs.n, self.device, self._gfs.dtype
            z = lambda k=n, d=wt: torch.zeros(k, dtype=d, device=dev)  # noqa: E731
            self._bufs = {"w0": z(), "g0": z(), "gl0": z(), "resid": z(), "comm": z(n + 2, self._comm_dtype())}
            self._bufs_version = ver
        return self._bufs

    def _comm_dtype(self):
        """APTS_D sums the subdomain steps in the dtype of the reference's coalesced buffer
        (= construction dtype, apts_d.py:129-141); APTS_P merges full parameters (model dtype)."""
        return self._gfs.dtype

    def _update_param_group(self):                       # apts_base.py:167-170
        for key in self.param_groups[0].keys():
            if key != "params" and hasattr(self, key):
                self.param_groups[0][key] = getattr(self, key)

    def _sync_attributes_from_param_group(self):         # apts_base.py:172-175
        for key, value in self.param_groups[0].items():
            if key != "params":
                setattr(self, key, value)

    def update_pytorch_lr(self) -> None:
        for g in self.param_groups:
            g["lr"] = self.delta

    def zero_grad(self, set_to_none: bool = False) -> None:
        self._gfs.zero_grad()

    def _push_ctl(self):
        cur = dict(delta=float(self.delta), min_delta=self.min_delta, max_delta=self.max_delta,
                   nu_dec=self.nu_dec, nu_inc=self.nu_inc, inc_factor=self.inc_factor,
                   dec_factor=self.dec_factor, tol=self.tol)
        diff = {k: v for k, v in cur.items() if self._ctl_pushed.get(k) != v}
        if diff:
            self._ctl.set(**diff)
            self._ctl_pushed.update(diff)

    # ------------------------------------------------------------------ flat vectors (API parity)
    @torch.no_grad()
    def glob_grad_to_vector(self):
        return self._gfs.flat_grad().clone()

    @torch.no_grad()
    def loc_grad_to_vector(self):
        return self._lfs.flat_grad().clone()

    @torch.no_grad()
    def glob_params_to_vector(self):
        self._gfs.ensure()
        return self._gfs.data.clone()

    @torch.no_grad()
    def loc_params_to_vector(self):
        self._lfs.ensure()
        return self._lfs.data.clone()

    # ------------------------------------------------------------------ closures
    def _labels(self):
        lab = self.labels
        return lab.long() if torch.is_tensor(lab) else lab

    def non_foc_loc_closure(self, compute_grad: bool = False):
        """apts_base.py:245-260."""
        if compute_grad and self._loc_eval is not None:
            loss = self._loc_eval(self.inputs, self.labels)
            self.loc_grad_evals += 1
            return loss
        self.loc_opt.zero_grad()
        loss = self.criterion(self.loc_model(self.inputs), self._labels())
        if compute_grad:
            loss.backward(retain_graph=bool(torch.is_tensor(self.inputs) and self.inputs.requires_grad))
            self.loc_grad_evals += 1
        return loss.detach()

    def foc_loc_closure(self, compute_grad: bool = False):
        """apts_base.py:262-292: loss + resid.(w_loc - w_glob) when the models differ.  The
        term changes the loss VALUE only (flattened copies carry no autograd history, Q7)."""
        loss = self.non_foc_loc_closure(compute_grad)
        b = self._buffers()
        loss_t = as_device_scalar(loss, self.device)
        out = torch.empty_like(loss_t)
        ctl = self._ctl
        ctl.lib.apts_foc(ctl.sp, ctl.wp, _lib.ptr(self._lfs.data), _lib.ptr(self._gfs.data), _lib.ptr(b["resid"]),
                         _lib.ptr(loss_t), _lib.ptr(out), _lib.dt(loss_t.dtype), self.tol, self._gfs.n,
                         _lib.dt(b["resid"].dtype), _lib.dt(self._gfs.dtype), _lib.stream())
        return out

    def glob_closure_main(self, compute_grad: bool = False):
        """apts_base.py:294-318 with the gradient and loss averaged by one in-place all-reduce
        of [g || loss] on the flat gradient buffer."""
        with_grad = torch.is_grad_enabled() or compute_grad
        if with_grad and self._glob_eval is not None:
            loss = self._glob_eval(self.inputs, self.labels)
            self.grad_evals += 1.0
        else:
            self.zero_grad()
            loss = self.criterion(self.model(self.inputs), self._labels())
            if with_grad:
                loss.backward()
                self.grad_evals += 1.0
            loss = loss.detach()
        if self._distributed:
            fb = self._gfs.fb
            if with_grad and not self._is_ddp:
                self._gfs.flat_grad()                      # make sure the views hold the gradient
                ext = fb.grad_ext[: fb.n + 1]
                ext[fb.n].copy_(loss)
                self._allreduce_mean(ext)
                loss = ext[fb.n].clone()
            else:
                loss = loss.clone()
                self._allreduce_mean(loss)
        return loss

    def _allreduce_mean(self, t):
        if dist.get_backend() == "nccl":
            all_reduce(t, dist.ReduceOp.AVG)
        else:
            all_reduce(t, dist.ReduceOp.SUM)
            t.div_(self.nr_models)

    # ------------------------------------------------------------------ inner loops
    def _grad_norm_of(self, optim, grad):
        # norms the optimizer's own kernels already reduced (no extra launch, no host sync)
        # (LSSR1_TR works in the 2-norm whatever the APTS norm and reports both; TR reports its own)
        if self.norm_type == math.inf and hasattr(optim, "last_grad_inf"):
            gn = optim.last_grad_inf
        else:
            gn = getattr(optim, "last_grad_norm", None)
        if gn is None:
            gn = float(grad.norm(p=self.norm_type))
        return gn

    def _step_loop(self, optim, max_iters, loss, grad, closure=None, closure_d=None):
        """apts_base.py:372-400."""
        if closure is None:
            raise ValueError("Closure must be provided for global/local optimization step in APTS.")
        for _ in range(max_iters):
            if isinstance(optim, (TR, LSSR1_TR)):
                loss, grad = optim.step(closure=closure, loss=loss, grad=grad)
            elif isinstance(optim, TRAdam):                  # apts_base.py:379-382
                loss = optim.step(closure=closure)
                grad = optim._flat_grad()
            else:                                            # apts_base.py:384-390 (ASNTR call form)
                loss, grad = optim.step(closure_main=closure, closure_d=closure_d, loss=loss, grad=grad)
            if getattr(optim, "device_control", False):
                continue        # converged steps are predicated no-ops on the device
            if hasattr(optim, "tol") and self._grad_norm_of(optim, grad) <= optim.tol:   # apts_base.py:397
                break
        return loss, grad

    def loc_steps(self, loss, grad):
        """apts_base.py:402-418; the rank-averaged evaluation counter rides in the step all-reduce."""
        return self._step_loop(self.loc_opt, self.max_loc_iters, loss, grad, closure=self.loc_closure,
                               closure_d=self.loc_closure_d)

    def glob_steps(self, loss, grad, closure=None, closure_d=None):
        return self._step_loop(self.glob_opt, self.max_glob_iters, loss, grad,
                               closure=self.glob_closure_main if closure is None else closure)

    # ------------------------------------------------------------------ control step
    def _control(self, buf, mode, scale, clampv, f0, pred_ptr, pred_dt, aux_ptr, bt=None, pre_applied=False):
        """apts_base.py:432-501 with `pred` supplied (APTS_D / APTS_P).  `buf` holds the combined
        step (mode 0) or the merged parameters (mode 1); `pre_applied`: the trial point w0 + step is
        already in place (fused peer-memory combine).  Returns (loss, grad, delta)."""
        b = self._buffers()
        lib, ctl = self._ctl.lib, self._ctl
        n, wt = self._gfs.n, _lib.dt(self._gfs.dtype)
        st = _lib.stream()
        self._push_ctl()
        if not pre_applied:
            lib.apts_trial(_lib.ptr(self._gfs.data), _lib.ptr(b["w0"]), _lib.ptr(buf), mode, float(scale),
                           float(clampv), n, wt, _lib.dt(self._ctor_dtype if bt is None else bt), st)
        trial = as_device_scalar(self.glob_closure_main(compute_grad=True), self.device)
        f0 = as_device_scalar(f0, self.device)
        if trial.dtype != f0.dtype:
            trial = trial.to(f0.dtype)
        G = self._gfs.flat_grad()
        out = torch.empty_like(f0)
        lib.apts_control(ctl.sp, ctl.wp, _lib.ptr(f0), _lib.ptr(trial), _lib.dt(f0.dtype), pred_ptr, pred_dt,
                         _lib.ptr(self._gfs.data), _lib.ptr(b["w0"]), _lib.ptr(b["g0"]), _lib.ptr(G),
                         _lib.ptr(out), aux_ptr, n, _lib.dt(b["g0"].dtype), wt, st)
        if self._device_control:
            return out, b["g0"], None, None, None
        return (out, b["g0"]) + self._read_control()

    def _read_control(self):
        # numerical failures of inner optimizers that run sync-free (their blocks are not read) ride
        # on the control block's report read
        inner = [(i, o) for i, o in ((0, self.glob_opt), (1, self.loc_opt)) if getattr(o, "device_control", False)]
        if inner:
            lay, lib, st = self._ctl.lay, self._ctl.lib, _lib.stream()
            src, dst = lay.off("err_seen"), lay.off("inner_err")
            for i, o in inner:
                lib.state_link(self._ctl.sp, dst + i, o._ds.sp, src, 1.0, -math.inf, math.inf, st)
        rep = self._ctl.read()
        bad = [rep.inner_err[i] for i, _ in inner if rep.inner_err[i]]
        if bad:
            for _, o in inner:
                o._ds.set(err_seen=0.0)
            DeviceState.raise_for(bad[0])
        self.delta = rep.delta
        self._ctl_pushed["delta"] = float(rep.delta)
        self.update_pytorch_lr()
        self.last_control = dict(pred=rep.pred, rho=rep.rho, accept=rep.accept != 0.0, delta=rep.delta)
        gn = None
        if rep.accept != 0.0:
            gn = rep.out_ginf if self.norm_type == math.inf else math.sqrt(rep.out_gg)
        return rep.delta, rep.aux, gn

    def control_step(self, step: torch.Tensor, pred: torch.Tensor | None = None, closure=None):
        """Public form of apts_base.py:432-501 (``pred`` must be supplied; the pred=None branch is
        only used by the out-of-scope APTS_IP)."""
        if pred is None or closure is not None:
            raise _lib.DD4MLError("dd4ml_b200 control_step: only the `pred`-supplied form is implemented")
        b = self._buffers()
        b["w0"].copy_(self.init_glob_flat)
        b["g0"].copy_(self.init_glob_grad)
        pred_t = as_device_scalar(pred, self.device)
        step = step.contiguous()
        if step.dtype not in (torch.float32, self._gfs.dtype):
            step = step.to(self._gfs.dtype)
        out, g, delta, _, _ = self._control(step, 0, 1.0, 0.0, self.init_glob_loss,
                                            _lib.ptr(pred_t), _lib.dt(pred_t.dtype), None, bt=step.dtype)
        if delta is None:               # device-control mode: the public form reports delta to the host
            delta, _, _ = self._read_control()
        return out, g, delta
