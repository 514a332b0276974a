"""OBS (orthonormal-basis SR1) trust-region sub-problem solver, drop-in name for
``dd4ml.solvers.obs.OBS`` (obs.py:26-254).

On the hot path the OBS computation is not a separate call: it runs in the one-CTA k-space
kernel behind ``tr_propose`` / ``lssr1_begin`` (csrc/kspace.h ``tr_solve``), fed by the Gram
blocks of the L-SR1 pairs.  This class keeps the reference's entry points:
``solve_tr_subproblem(g, delta, gamma, Psi, Minv)`` and ``ComputeSBySMW(tau, g, PsiTg, Psi, Minv,
PsiPsi)`` accept either the owning ``LSR1`` in place of ``Psi`` (its Psi / Minv are implicit) or the
explicit (n x k) ``Psi`` and (k x k) ``Minv`` tensors of the reference signature (obs.py:46,175):
the explicit form runs the same kernels on a scratch L-SR1 block with S = 0, Y = Psi^T, so that
Y - gamma S = Psi, and the caller's Minv in the state block (``use_mx``).
"""
from __future__ import annotations

import torch

from .. import _lib


class OBS:
    def __init__(self):
        self.tol = 1e-6          # obs.py:29
        self.last: dict = {}     # diagnostics of the last solve (case, sigma*, Newton iterations)

    @staticmethod
    def _explicit(Psi: torch.Tensor, Minv: torch.Tensor, gamma):
        """Scratch L-SR1 block that represents the caller's explicit (Psi, Minv, gamma)."""
        from ..optimizers.lsr1 import LSR1
        if not (torch.is_tensor(Psi) and Psi.dim() == 2 and torch.is_tensor(Minv) and Minv.dim() == 2):
            raise TypeError("dd4ml_b200.OBS: `Psi` must be an LSR1 object or an (n x k) tensor with a (k x k) `Minv`")
        n, k = Psi.shape
        if Minv.shape != (k, k):
            raise ValueError(f"Minv must be ({k} x {k}) for an (n x {k}) Psi")
        if not Psi.is_cuda:
            raise _lib.DD4MLError("dd4ml_b200 has no CPU path: Psi must live on a CUDA device")
        dtype = Psi.dtype if Psi.dtype in (torch.float32, torch.float64) else torch.float32
        h = LSR1(gamma=float(gamma), memory_length=max(k, 1), tol=1e-6, device=Psi.device, dtype=dtype)
        if k > h.ds.lay.maxm:
            raise _lib.DD4MLError(f"dd4ml_b200.OBS: k = {k} exceeds the build's maximum {h.ds.lay.maxm}")
        h.ensure_storage(n)
        lay, ms, mm = h.ds.lay, h.ds.lay.maxs, h.ds.lay.maxm
        h._Ybuf[:k, :n].copy_(Psi.t())                        # S stays 0:  Y - gamma S = Psi
        P64 = Psi.to(torch.float64)
        yy = torch.zeros(ms, ms, dtype=torch.float64, device=Psi.device)
        yy[:k, :k] = P64.t() @ P64                             # the k x k Gram block the kernels keep incrementally
        mx = torch.zeros(mm, mm, dtype=torch.float64, device=Psi.device)
        mx[:k, :k] = Minv.to(torch.float64)
        st = h.ds.state
        st[lay.slice("YY")].copy_(yy.reshape(-1))
        st[lay.slice("Mx")].copy_(mx.reshape(-1))
        st[lay.slice("order")].copy_(torch.arange(mm, dtype=torch.float64, device=Psi.device))
        h.ds.set(k=k, spare=k, gamma=float(gamma), use_mx=1.0, second_order=1.0)
        h._k_host, h._gamma_host = k, float(gamma)
        return h

    def solve_tr_subproblem(self, g, delta, gamma, Psi, Minv=None):
        from ..optimizers.lsr1 import LSR1
        from ..utility.optimizer_utils import solve_tr_second_order
        for name, t in (("g", g), ("delta", torch.as_tensor(delta)), ("gamma", torch.as_tensor(gamma))):
            if torch.isnan(t).any() or torch.isinf(t).any():
                raise ValueError(f"{name} contains NaN or Inf values")
        if float(delta) < 0:
            raise ValueError("Delta must be non-negative")
        if not isinstance(Psi, LSR1):
            for name, t in (("Psi", Psi), ("Minv", Minv)):                    # obs.py:54-57
                if torch.is_tensor(t) and (torch.isnan(t).any() or torch.isinf(t).any()):
                    raise ValueError(f"{name} contains NaN or Inf values")
            h = self._explicit(Psi, Minv, gamma)
            # (the reference's OBS has no ||g|| <= tol exit: solve whatever the gradient norm)
            p, _ = solve_tr_second_order(g, torch.norm(g), float(delta), h, self, 0.0, dogleg=False)
            return p
        h = Psi
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
            Psi = self._explicit(Psi, Minv, 1.0)      # gamma does not enter the SMW formula (obs.py:175-182)
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
