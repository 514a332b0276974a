"""TRAdam -- Adam direction clipped to a trust radius (reference: optimizers/tradam.py:7-111).

Same constructor and ``step(closure)`` as the reference class; the 12 ATen launches of a step
(moment updates, bias correction, direction, norm, scale, per-tensor scatter) are two fused
kernels (csrc/k_tradam.cu) over the flat parameter / gradient buffers, and the ``.item()`` of
the step length (tradam.py:99-101) stays on the device: a step performs no host read.

Standalone optimizer only.  Under APTS the reference drives it with closures that do not
back-propagate (apts_base.py:379-382), which makes the local phase a no-op there; the APTS
classes of this package refuse ``loc_opt='tradam'`` rather than reproduce that.
"""
from __future__ import annotations

from typing import Callable, Iterable

import torch
from torch.optim import Optimizer

from .. import _lib
from .._device import DeviceState
from ..flat import get_flat


class TRAdam(Optimizer):
    __name__ = "TRAdam"

    def __init__(self, params: Iterable[torch.nn.Parameter], *,
                 flat_grads_fn: Callable[[], torch.Tensor] | None = None,
                 flat_params_fn: Callable[[], torch.Tensor] | None = None, **kwargs) -> None:
        self.lr = float(kwargs.pop("lr"))
        self.betas = kwargs.pop("betas", (0.9, 0.999))
        self.eps = float(kwargs.pop("eps", 1e-8))
        self.norm_type = kwargs.pop("norm_type", torch.inf)
        if self.norm_type not in (2, torch.inf):
            raise ValueError("norm_type must be 2 or torch.inf")
        self._flat_grads_fn = flat_grads_fn
        self._flat_params_fn = flat_params_fn
        super().__init__(params, {"lr": self.lr})
        self.ps = [p for g in self.param_groups for p in g["params"]]
        self.shapes = [p.shape for p in self.ps]
        self.numels = [p.numel() for p in self.ps]
        self._fs = get_flat(self.ps)
        device, n = self._fs.device, self._fs.n
        self._ds = DeviceState(device)
        # tradam.py:45-48: moment / step buffers in the default dtype (float32)
        self._grad_buf = torch.zeros(n, device=device)
        self._step_buf = torch.zeros_like(self._grad_buf)
        self._m_buf = torch.zeros_like(self._grad_buf)
        self._v_buf = torch.zeros_like(self._grad_buf)
        self.t = 0

    def reset_momentum(self):
        """tradam.py:52-55."""
        self._m_buf.zero_()
        self._v_buf.zero_()

    def zero_grad(self, set_to_none: bool = False) -> None:
        self._fs.zero_grad()

    def _flat_grad(self) -> torch.Tensor:
        """tradam.py:57-66: the current gradient as a fresh flat vector (buffer dtype)."""
        if self._flat_grads_fn is not None:
            g = self._flat_grads_fn()
            return g.clone() if isinstance(g, torch.Tensor) else g
        self._grad_buf.copy_(self._fs.flat_grad())
        return self._grad_buf.clone()

    def _apply_update(self, sign: float = 1.0) -> None:
        """tradam.py:68-78 (compatibility; ``step`` applies the update inside the kernel)."""
        with torch.no_grad():
            if self._flat_params_fn is not None:
                updated = self._flat_params_fn() + sign * self._step_buf
                torch.nn.utils.vector_to_parameters(updated, self.ps)
            else:
                self._fs.data.add_(self._step_buf.to(self._fs.dtype), alpha=sign)

    @property
    def step_length(self) -> float:
        """Length of the unclipped Adam direction of the last step (one host read)."""
        return float(self._ds.read().pnorm)

    def step(self, closure=None):
        lib, ds = self._ds.lib, self._ds
        self.t += 1
        loss = closure() if closure is not None else None
        if self._flat_grads_fn is not None:
            g = self._flat_grads_fn().detach()
            if g.dtype != self._m_buf.dtype:
                g = g.to(self._m_buf.dtype)
            g = g.contiguous()
        else:
            G = self._fs.flat_grad()
            if G.dtype != self._m_buf.dtype:
                self._grad_buf.copy_(G)
                G = self._grad_buf
            g = G
        n = self._m_buf.numel()
        b1, b2 = float(self.betas[0]), float(self.betas[1])
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        vt = _lib.dt(self._m_buf.dtype)
        s = _lib.stream()
        lib.tradam_moments(ds.sp, ds.wp, _lib.ptr(g), _lib.ptr(self._m_buf), _lib.ptr(self._v_buf), b1, b2, bc1, bc2,
                           self.eps, self.lr, 1 if self.norm_type == torch.inf else 0, n, vt, s)
        if self._flat_params_fn is not None:
            # tradam.py:71-74: parameters supplied by the caller: update a copy and scatter it
            w = self._flat_params_fn().detach().clone().contiguous()
            lib.tradam_apply(ds.sp, ds.wp, _lib.ptr(self._m_buf), _lib.ptr(self._v_buf), _lib.ptr(w),
                             _lib.ptr(self._step_buf), bc1, bc2, self.eps, n, vt, _lib.dt(w.dtype), s)
            with torch.no_grad():
                torch.nn.utils.vector_to_parameters(w, self.ps)
        else:
            self._fs.ensure()
            w = self._fs.data
            lib.tradam_apply(ds.sp, ds.wp, _lib.ptr(self._m_buf), _lib.ptr(self._v_buf), _lib.ptr(w),
                             _lib.ptr(self._step_buf), bc1, bc2, self.eps, n, vt, _lib.dt(w.dtype), s)
        return loss

    # ------------------------------------------------------------------ checkpointing
    def state_dict(self):
        sd = super().state_dict()
        sd["dd4ml_b200"] = {"t": self.t, "m": self._m_buf.clone(), "v": self._v_buf.clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        extra = state_dict.pop("dd4ml_b200", None)
        super().load_state_dict(state_dict)
        if extra is not None:
            self.t = int(extra["t"])
            self._m_buf.copy_(extra["m"])
            self._v_buf.copy_(extra["v"])
