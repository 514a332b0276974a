"""APTS with parameter subdomains, drop-in for ``dd4ml.optimizers.apts_p.APTS_P``
(apts_p.py:4-245).

Each rank trains a random disjoint subset of parameter TENSORS (apts_utils.py:83-124).  The
local clone is laid out with the owned tensors first, so the local optimizer works on a
contiguous prefix of the local flat buffer (the reference concatenates the trainable tensors
with ``torch.cat`` on every evaluation, apts_p.py:156-165).  The merge of the subdomain
updates is ONE all-reduce of a zero-filled flat buffer (the reference: T per-tensor
all-reduces, apts_p.py:145-153) with the summed predicted reduction appended; the
element-wise clamp to the infinity-norm radius (apts_p.py:221-224) is fused into the kernel
that applies the step.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.distributed as dist

from .. import _lib
from ..flat import FlatBuffer, FlatSlice
from ..utility.apts_utils import clone_model, mark_trainable
from . import apts_base
from .apts_base import APTS_Base
from .tr import as_device_scalar


class APTS_P(APTS_Base):
    __name__ = "APTS_P"

    @staticmethod
    def setup_APTS_hparams(config):
        return APTS_Base.setup_APTS_hparams(config)

    def __init__(self, params, model=None, delta=None, min_delta=None, max_delta=None, nu_dec=None,
                 nu_inc=None, inc_factor=None, dec_factor=None, criterion=None, device=None,
                 nr_models=None, glob_opt=None, glob_opt_hparams=None, loc_opt=None, loc_opt_hparams=None,
                 *, glob_pass=True, norm_type=math.inf, max_loc_iters=3, max_glob_iters=3, tol=1e-6):
        super().__init__(params, model=model, delta=delta, min_delta=min_delta, max_delta=max_delta,
                         nu_dec=nu_dec, nu_inc=nu_inc, inc_factor=inc_factor, dec_factor=dec_factor,
                         criterion=criterion, device=device, nr_models=nr_models, glob_opt=glob_opt,
                         glob_opt_hparams=glob_opt_hparams, loc_opt=loc_opt, loc_opt_hparams=loc_opt_hparams,
                         glob_pass=glob_pass, norm_type=math.inf, max_loc_iters=max_loc_iters,
                         max_glob_iters=max_glob_iters, tol=tol)
        self.loc_model = mark_trainable(clone_model(model))
        lps = list(self.loc_model.parameters())
        owned_idx = [i for i, p in enumerate(lps) if p.requires_grad]
        other_idx = [i for i, p in enumerate(lps) if not p.requires_grad]
        self._owned_idx = owned_idx
        # local layout: owned tensors (model order) first, then the rest
        fb = FlatBuffer([lps[i] for i in owned_idx + other_idx])
        self._lfb = fb
        self._lfs = FlatSlice(fb, 0, fb.n)
        owned = [lps[i] for i in owned_idx]
        from .lssr1_tr import LSSR1_TR
        from .tr import TR
        if issubclass(loc_opt, (TR, LSSR1_TR)):
            self.loc_opt = loc_opt(params=owned, **loc_opt_hparams)
        else:
            # apts_p.py:76-81 hands every local optimizer the flat hooks: TRAdam takes them, a stock
            # torch optimizer (loc_opt: sgd) raises TypeError here, as in the reference
            self.loc_opt = loc_opt(params=owned, flat_grads_fn=self.loc_grad_to_vector,
                                   flat_params_fn=self.loc_params_to_vector, **loc_opt_hparams)
        # the reference passes flat_grads_fn / flat_params_fn hooks to this optimizer (apts_p.py:76-81):
        # its vectors bypass the construction-dtype flat buffers of TR / LSSR1_TR
        self.loc_opt._ref_hooks = True
        self.loc_closure = self.non_foc_loc_closure
        self._layout()
        self.sqrt_n_global = math.sqrt(self.n_global)
        self.sqrt_n_local = math.sqrt(max(self.n_local, 1))
        # apts_p.py:92-104: infinity-norm radii -> 2-norm radii of the inner optimizers
        self.glob_opt.max_delta = self.max_delta * self.sqrt_n_global
        self.glob_opt.min_delta = self.min_delta * self.sqrt_n_global
        self.loc_opt.max_delta = self.glob_opt.max_delta
        self.loc_opt.min_delta = self.glob_opt.min_delta
        self.glob_opt.delta = max(self.glob_opt.min_delta, min(self.glob_opt.max_delta, self.delta * self.sqrt_n_global))
        self.loc_opt.delta = self.glob_opt.delta
        self._has_buffers = any(True for _ in self.loc_model.buffers())
        self._enable_device_control()
        from ..graphs import GraphedEval
        self._loc_eval = GraphedEval(self.loc_model, self.criterion, self._lfs)

    def _layout(self):
        """Segment tables global offset <-> local offset of every parameter tensor, and the sizes
        derived from them.  Offsets are in elements, so they survive a dtype change; they are
        rebuilt together with the flat buffers all the same (`_rebind`)."""
        fb = self._lfb
        lps = list(self.loc_model.parameters())
        owned_idx = self._owned_idx
        other_idx = [i for i in range(len(lps)) if i not in set(owned_idx)]
        self.n_local_owned = sum(lps[i].numel() for i in owned_idx)
        self.n_global = self._gfs.n
        self.n_local = self.n_local_owned
        goff = self._gfs.fb.offsets
        order = owned_idx + other_idx
        loff = {t: fb.offsets[j] for j, t in enumerate(order)}
        numel = [p.numel() for p in lps]
        mk = lambda xs: (ctypes.c_int64 * len(xs))(*xs)  # noqa: E731
        self._seg_owned = (len(owned_idx), mk([goff[t] for t in owned_idx]), mk([loff[t] for t in owned_idx]),
                           mk([numel[t] for t in owned_idx]))
        self._seg_all = (len(order), mk([loff[t] for t in order]), mk([goff[t] for t in order]),
                         mk([numel[t] for t in order]))

    def _rebind(self):
        """The reference Trainer converts ``model`` and ``optimizer.loc_model`` to float64 AFTER the
        optimizer exists (trainer.py:1115-1128, 1203-1211), which replaces every parameter's storage.
        Both flat buffers are re-bound to the new storages (FlatBuffer.ensure), the segment tables and
        the n-vector work buffers are rebuilt, and the inner optimizers pick the new buffers up
        through their FlatSlice versions -- as APTS_D does."""
        g_re = self._gfs.ensure()
        l_re = self._lfb.ensure()
        if not (g_re or l_re):
            return False
        if self._gfs.dtype != self._lfb.dtype:
            raise _lib.DD4MLError(
                f"dd4ml_b200 APTS_P: model is {self._gfs.dtype} but loc_model is {self._lfb.dtype}; convert both "
                "(the reference Trainer does: trainer.py:1203-1211)")
        self._layout()
        self._buffers()
        # the clone mirrors the global model between outer iterations: re-establish it in the new storage
        cnt, dsto, srco, lens = self._seg_all
        self._ctl.lib.seg_copy(_lib.ptr(self._lfb.data), _lib.ptr(self._gfs.data), cnt, dsto, srco, lens,
                               self._lfb.n, 0, _lib.dt(self._gfs.dtype), _lib.stream())
        return True

    # ------------------------------------------------------------------ flat vectors (API parity)
    @torch.no_grad()
    def loc_params_to_vector(self):
        return self._lfb.data[: self.n_local_owned].clone()

    @torch.no_grad()
    def loc_grad_to_vector(self):
        return self._lfb.flat_grad()[: self.n_local_owned].clone()

    # ------------------------------------------------------------------ apts_p.py:115-153
    @torch.no_grad()
    def sync_loc_to_glob(self, comm=None, pred=None, evals=None):
        """Merge the owned updates of all ranks: comm[0:n] <- owned slices (zero elsewhere),
        comm[n] <- this rank's predicted reduction, comm[n+1] <- its evaluation counter,
        one all-reduce(SUM)."""
        b = self._buffers()
        comm = b["comm"] if comm is None else comm
        lib = self._ctl.lib
        cnt, dsto, srco, lens = self._seg_owned
        lib.seg_copy(_lib.ptr(comm), _lib.ptr(self._lfb.data), cnt, dsto, srco, lens, self.n_global + 2, 1,
                     _lib.dt(comm.dtype), _lib.stream())
        if pred is not None:
            comm[self.n_global].copy_(pred)
        if evals is not None:
            comm[self.n_global + 1:self.n_global + 2].fill_(float(evals))   # (fill_, not item assignment: no host->device copy)
        if self._distributed:
            apts_base.all_reduce(comm, dist.ReduceOp.SUM)
        return comm

    # ------------------------------------------------------------------ apts_p.py:167-180
    @torch.no_grad()
    def sync_glob_to_loc(self):
        if self._device_control:
            self._link(self._ctl, self.glob_opt._ds, 1.0 / self.sqrt_n_global, self.min_delta, self.max_delta)
            self._link(self.loc_opt._ds, self.glob_opt._ds)
            self.loc_opt._delta_stale = True
        else:
            self.delta = min(self.max_delta, max(self.min_delta, self.glob_opt.delta / self.sqrt_n_global))
            self.update_pytorch_lr()
            self.loc_opt.delta = self.glob_opt.delta
            if hasattr(self.loc_opt, "update_pytorch_lr"):
                self.loc_opt.update_pytorch_lr()
        cnt, dsto, srco, lens = self._seg_all
        self._ctl.lib.seg_copy(_lib.ptr(self._lfb.data), _lib.ptr(self._gfs.data), cnt, dsto, srco, lens,
                               self._lfb.n, 0, _lib.dt(self._gfs.dtype), _lib.stream())
        if self._has_buffers:
            base = self.model.module if hasattr(self.model, "module") else self.model
            for lb, gb in zip(self.loc_model.buffers(), base.buffers()):
                lb.copy_(gb)

    # ------------------------------------------------------------------ apts_p.py:182-245
    def step(self, inputs, labels, inputs_d=None, labels_d=None, hNk=None):
        lib = self._ctl.lib
        self.inputs, self.labels = inputs, labels
        self.inputs_d, self.labels_d, self.hNk = inputs_d, labels_d, hNk
        self.grad_evals, self.loc_grad_evals = 0.0, 0.0
        if self._resync:                 # first step after load_state_dict
            self.sync_glob_to_loc()
            self._resync = False
        self._rebind()
        b = self._buffers()
        n = self.n_global
        self.glob_opt._known_gn = self.loc_opt._known_gn = None
        b["w0"].copy_(self._gfs.data)
        self.init_glob_flat = b["w0"]
        self.init_glob_loss = self.glob_closure_main(compute_grad=True)
        b["g0"].copy_(self._gfs.flat_grad())
        self.init_glob_grad = b["g0"]
        self.init_loc_loss = as_device_scalar(self.loc_closure(compute_grad=True), self.device)
        gl0 = b["gl0"][: self.n_local_owned]
        gl0.copy_(self._lfb.flat_grad()[: self.n_local_owned])
        self.init_loc_grad = gl0

        loc_loss, _ = self.loc_steps(self.init_loc_loss, self.init_loc_grad)
        loc_loss = as_device_scalar(loc_loss, self.device, self.init_loc_loss.dtype)

        comm = b["comm"]
        pred_loc = (self.init_loc_loss - loc_loss).to(comm.dtype)
        self.sync_loc_to_glob(comm, pred_loc, self.loc_grad_evals)
        es = comm.element_size()
        loss, grad, new_delta, evals_sum, gn = self._control(
            comm, 1, 1.0, self.delta, self.init_glob_loss, comm.data_ptr() + n * es, _lib.dt(comm.dtype),
            comm.data_ptr() + (n + 1) * es)
        if self._device_control:
            self._link(self.glob_opt._ds, self._ctl, self.sqrt_n_global, hi=self.glob_opt.max_delta)
            self.glob_opt._delta_stale = True
            if self.glob_pass:
                loss, grad = self.glob_steps(loss, grad)
            self.sync_glob_to_loc()
            new_delta, evals_sum, gn = self._read_control()           # the only host read of the iteration
        else:
            self.delta = new_delta
            self.glob_opt.delta = min(self.glob_opt.max_delta, new_delta * self.sqrt_n_global)
            if self.glob_pass:
                self.glob_opt._known_gn = None if gn is None else (id(grad), grad._version, grad.data_ptr(), gn)
                loss, grad = self.glob_steps(loss, grad)
            self.sync_glob_to_loc()
        self.grad_evals += evals_sum / (self.nr_models if self._distributed else 1)
        return loss
