"""OBS (orthonormal-basis SR1) trust-region sub-problem solver, drop-in name for
``dd4ml.solvers.obs.OBS`` (obs.py:26-254).

On the hot path the OBS computation is not a separate call: it runs in k-space inside the
last CTA of ``tr_propose`` / ``lssr1_begin`` (csrc/kspace.h ``tr_solve``), fed by the Gram
blocks of the L-SR1 pairs.  This class keeps the reference's entry point for callers that
hold an ``LSR1`` object: ``solve_tr_subproblem(g, delta, gamma, Psi, Minv)`` accepts the
owning ``LSR1`` in place of ``Psi`` (its Psi / Minv are implicit) and returns p*.
"""
from __future__ import annotations

import torch

from .. import _lib


class OBS:
    def __init__(self):
        self.tol = 1e-6          # obs.py:29
        self.last: dict = {}     # diagnostics of the last solve (case, sigma*, Newton iterations)

    def solve_tr_subproblem(self, g, delta, gamma, Psi, Minv=None):
        from ..optimizers.lsr1 import LSR1
        from ..utility.optimizer_utils import solve_tr_second_order
        if not isinstance(Psi, LSR1):
            raise TypeError(
                "dd4ml_b200.OBS works on the device-resident L-SR1 state: pass the LSR1 object as `Psi` "
                "(explicit (n x k) Psi / Minv matrices are never materialised in this implementation)")
        h = Psi
        for name, t in (("g", g), ("delta", torch.as_tensor(delta)), ("gamma", torch.as_tensor(gamma))):
            if torch.isnan(t).any() or torch.isinf(t).any():
                raise ValueError(f"{name} contains NaN or Inf values")
        if float(delta) < 0:
            raise ValueError("Delta must be non-negative")
        dog = h.ds.read().dogleg
        p, _ = solve_tr_second_order(g, torch.norm(g), float(delta), h, self, h.tol, dogleg=False)
        h.ds.set(dogleg=dog)
        return p

    # ------------------------------------------------------------------ obs.py:175-254 as stand-alone calls
    def _secular(self, mode, x0, Lambda, a_j, delta):
        dev = a_j.device if (torch.is_tensor(a_j) and a_j.is_cuda) else torch.device("cuda")
        dtype = a_j.dtype if torch.is_tensor(a_j) and a_j.dtype in (torch.float32, torch.float64) else torch.float32
        L = torch.as_tensor(Lambda, dtype=dtype).to(dev).contiguous().reshape(-1)
        A = torch.as_tensor(a_j, dtype=dtype).to(dev).contiguous().reshape(-1)
        if L.numel() != A.numel():
            raise ValueError("Lambda and a_j must have the same length")
        out = torch.empty(3, dtype=torch.float64, device=dev)
        lib = _lib.get()
        with torch.cuda.device(dev):
            lib.obs_secular(mode, float(x0), _lib.ptr(L), _lib.ptr(A), int(A.numel()), float(delta), self.tol,
                            _lib.dt(dtype), _lib.ptr(out), _lib.stream())
        return out, dtype

    def phiBar_fg(self, sigma, Dd, a_j, delta):
        """obs.py:227-254: (phiBar, phiBar') of the secular equation, degenerate branch included."""
        out, dtype = self._secular(1, sigma, Dd, a_j, delta)
        return out[0].to(dtype), out[1].to(dtype)

    def phiBar_f(self, sigma, Dd, a_j, delta):
        """obs.py:184-208."""
        out, dtype = self._secular(0, sigma, Dd, a_j, delta)
        return out[0].to(dtype)

    def Newton(self, x0, Lambda, a_j, delta):
        """obs.py:210-225: sigma <- sigma - phiBar/phiBar' from x0, at most 200 iterations."""
        out, dtype = self._secular(2, x0, Lambda, a_j, delta)
        return out[0].to(dtype)

    def ComputeSBySMW(self, tauStar, g, PsiTg=None, Psi=None, Minv=None, PsiPsi=None):
        """obs.py:175-182: p = -g/tau + Psi ((tau Minv + Psi^T Psi) \\ Psi^T g).  ``Psi`` is the owning
        ``LSR1`` (its Psi, Minv and Psi^T Psi are implicit; PsiTg is recomputed in the same pass)."""
        from .._device import DeviceState
        from ..optimizers.lsr1 import LSR1
        if not isinstance(Psi, LSR1):
            raise TypeError("dd4ml_b200.OBS.ComputeSBySMW: pass the LSR1 object as `Psi`")
        h, ds = Psi, Psi.ds
        g = g.detach().contiguous()
        n, vt = g.numel(), _lib.dt(g.dtype)
        h.ensure_storage(n)
        ds.set(tau=float(tauStar))
        p = torch.empty_like(g)
        ds.lib.tr_propose(ds.sp, ds.wp, _lib.ptr(g), h.S_ptr, h.Y_ptr, n, h.ld, vt, _lib.dt(h.dtype), 3,
                          h.memory_length, _lib.stream())
        ds.lib.tr_apply(ds.sp, ds.wp, _lib.ptr(g), h.S_ptr, h.Y_ptr, _lib.ptr(p), None, n, h.ld, vt,
                        _lib.dt(h.dtype), vt, h.memory_length, _lib.stream())
        DeviceState.raise_for(ds.read().err)
        return p
