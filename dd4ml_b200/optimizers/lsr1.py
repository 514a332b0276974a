"""Compact limited-memory SR1 on the device (drop-in for ``dd4ml.optimizers.lsr1.LSR1``).

    B = gamma I + Psi M Psi^T,  Psi = Y - gamma S,
    M^{-1} = diag(S^T Y) + L + L^T - gamma S^T S + tol ||.||_F I      (lsr1.py:146-179)

Layout: the pairs live pair-major in two ring buffers of MAXS physical
slots x ld elements (each s_i / y_i contiguous; one spare slot receives the candidate
pair), instead of the reference's (n x m) row-major matrices whose column shift costs
2 n (m-1) reads+writes per accepted pair (lsr1.py:122-127).  Psi and M^{-1} are never
materialised in n-space: the k x k Gram blocks S^T S, S^T Y, Y^T Y are kept in the
device state block and everything the reference derives from Psi (precompute, B(v)
inside update_memory, OBS) is evaluated from them in the k-space kernel behind the fused
reduction kernels (csrc/kspace.h).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib
from .._device import DeviceState


class _PairList:
    """Stand-in for the reference's ``_S`` / ``_Y`` python lists: ``len()`` is the number of
    stored pairs (apts_base.py:466-473, tr.py:151 only use the length)."""

    def __init__(self, owner, which):
        self._o, self._w = owner, which

    def __len__(self):
        return self._o.num_pairs()

    def __getitem__(self, i):
        k = self._o.num_pairs()
        if i < 0:
            i += k
        if not 0 <= i < k:
            raise IndexError(i)
        order = self._o._order()
        buf = self._o._Sbuf if self._w == "S" else self._o._Ybuf
        return buf[order[i], :self._o.n]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class LSR1:
    def __init__(self, gamma: float = 1.0, memory_length: int = 10, tol: float = 1e-6,
                 device: Optional[torch.device] = None, dtype: torch.dtype | None = None):
        lib = _lib.get()
        self.memory_length = int(memory_length)
        if self.memory_length > lib.max_mem_length:
            raise _lib.DD4MLError(
                f"dd4ml_b200: memory_length={memory_length} exceeds this build's maximum "
                f"{lib.max_mem_length} (DD4ML_MAXM in csrc/state.h)")
        self.tol = float(tol)
        self.device = torch.device("cuda" if device is None else device)
        if self.device.type != "cuda":
            raise _lib.DD4MLError("dd4ml_b200 has no CPU path: LSR1 needs a CUDA device")
        self.dtype = torch.float32 if dtype is None else dtype
        self.ds = DeviceState(self.device, gamma0=float(gamma))
        self.ds.set(mem_len=self.memory_length, tol=self.tol, second_order=1.0)
        self._gamma0 = float(gamma)
        self.n = None
        self.ld = 0
        self._Sbuf = self._Ybuf = None
        self._k_host = 0            # mirrors, refreshed from every report the owner reads
        self._gamma_host = float(gamma)
        self._S = _PairList(self, "S")
        self._Y = _PairList(self, "Y")

    # ------------------------------------------------------------------ storage
    def ensure_storage(self, n: int):
        if self._Sbuf is not None and self.n == n:
            return
        if self._Sbuf is not None and self._k_host > 0:
            raise _lib.DD4MLError("dd4ml_b200: LSR1 vector length changed while pairs are stored")
        self.n = int(n)
        self.ld = (self.n + 3) // 4 * 4
        slots = self.ds.lay.maxs
        self._Sbuf = torch.zeros(slots, self.ld, dtype=self.dtype, device=self.device)
        self._Ybuf = torch.zeros(slots, self.ld, dtype=self.dtype, device=self.device)

    @property
    def S_ptr(self):
        return self._Sbuf.data_ptr()

    @property
    def Y_ptr(self):
        return self._Ybuf.data_ptr()

    # ------------------------------------------------------------------ host mirrors
    def absorb(self, rep):
        """Refresh the host mirrors from a report the owner just read."""
        self._k_host = int(rep.k)
        self._gamma_host = float(rep.gamma)
        self._stale = False

    def num_pairs(self) -> int:
        if getattr(self, "_stale", False):      # device-control mode: mirrors are refreshed on demand
            self.absorb(self.ds.read())
        return self._k_host

    def _order(self):
        rep = self.ds.read(full=True)
        self.absorb(rep)
        return [int(x) for x in rep.order[: self._k_host]]

    @property
    def gamma(self) -> torch.Tensor:
        if getattr(self, "_stale", False):
            self.absorb(self.ds.read())
        return torch.tensor(self._gamma_host, dtype=self.dtype, device=self.device)

    @gamma.setter
    def gamma(self, value):
        self._gamma_host = float(value)
        self.ds.set(gamma=self._gamma_host)

    @property
    def _current_pairs(self) -> int:
        return self._k_host

    # ------------------------------------------------------------------ reference API
    def update_memory(self, s: torch.Tensor, y: torch.Tensor) -> None:
        """lsr1.py:53-144 as one fused pass + k-space acceptance chain."""
        s = s.detach().flatten()
        y = y.detach().flatten()
        if s.dtype != y.dtype:
            y = y.to(s.dtype)
        self.ensure_storage(s.numel())
        lib = self.ds.lib
        lib.lsr1_update(self.ds.sp, self.ds.wp, _lib.ptr(s.contiguous()), _lib.ptr(y.contiguous()),
                        self.S_ptr, self.Y_ptr, self.n, self.ld, _lib.dt(s.dtype), _lib.dt(self.dtype),
                        self.memory_length, _lib.stream())
        rep = self.ds.read()
        DeviceState.raise_for(rep.err)
        self.absorb(rep)

    @torch.no_grad()
    def set_memory(self, S: torch.Tensor, Y: torch.Tensor, gamma: float) -> None:
        """Install k pairs (columns of the (n x k) matrices S, Y, oldest first) and gamma directly, bypassing
        the acceptance chain -- what writing the reference's `_S_matrix` / `_Y_matrix` / `gamma` attributes
        does (lsr1.py:112-120).  Compatibility / test utility (library GEMMs for the Gram blocks), not on
        the hot path."""
        S, Y = S.detach().to(self.device), Y.detach().to(self.device)
        n, k = S.shape
        if k > self.memory_length:
            raise ValueError(f"{k} pairs exceed memory_length={self.memory_length}")
        self._Sbuf = None
        self._k_host = 0
        self.ensure_storage(n)
        lay, ms = self.ds.lay, self.ds.lay.maxs
        self._Sbuf[:k, :n].copy_(S.t())
        self._Ybuf[:k, :n].copy_(Y.t())
        S64, Y64 = self._Sbuf[:k, :n].double(), self._Ybuf[:k, :n].double()      # (values as stored)
        st = self.ds.state
        for name, A, B in (("SS", S64, S64), ("SY", S64, Y64), ("YY", Y64, Y64)):
            blk = torch.zeros(ms, ms, dtype=torch.float64, device=self.device)
            blk[:k, :k] = A @ B.t()
            st[lay.slice(name)].copy_(blk.reshape(-1))
        st[lay.slice("order")].copy_(torch.arange(lay.maxm, dtype=torch.float64, device=self.device))
        self.ds.set(k=k, spare=k, gamma=float(gamma), use_mx=0.0)
        self._k_host, self._gamma_host = k, float(gamma)

    def precompute(self) -> None:
        """lsr1.py:146-179.  Nothing to do: Psi / M^{-1} are k-space quantities rebuilt from
        the Gram blocks inside the kernels.  Kept for API compatibility."""
        return None

    def B(self, v: torch.Tensor) -> torch.Tensor:
        """lsr1.py:181-198: gamma v + Psi (Minv \\ Psi^T v); two fused launches."""
        v = v.detach().flatten().contiguous()
        if v.dtype not in (torch.float32, torch.float64) or (self.dtype == torch.float64 and v.dtype != torch.float64):
            v = v.to(self.dtype)
        out = torch.empty_like(v)
        if self._Sbuf is None:
            self.ensure_storage(v.numel())
        lib = self.ds.lib
        lib.lsr1_B(self.ds.sp, self.ds.wp, _lib.ptr(v), self.S_ptr, self.Y_ptr, _lib.ptr(out), v.numel(),
                   self.ld, _lib.dt(v.dtype), _lib.dt(self.dtype), self.memory_length, _lib.stream())
        return out.to(self.dtype)

    # Debug / compatibility views (host-synchronising; not used on the hot path) ------
    @property
    def S(self) -> torch.Tensor:
        order = self._order()
        if not order:
            return torch.zeros((0, 0), device=self.device, dtype=self.dtype)
        return self._Sbuf[order, : self.n].t()

    @property
    def Y(self) -> torch.Tensor:
        order = self._order()
        if not order:
            return torch.zeros((0, 0), device=self.device, dtype=self.dtype)
        return self._Ybuf[order, : self.n].t()

    @property
    def Psi(self) -> torch.Tensor:
        if self._k_host == 0:
            return torch.zeros((0, 0), device=self.device, dtype=self.dtype)
        return self.Y - self.gamma * self.S

    @property
    def Minv(self) -> torch.Tensor:
        rep = self.ds.read(full=True)
        self.absorb(rep)
        k, ms = self._k_host, self.ds.lay.maxs
        if k == 0:
            return torch.zeros((0, 0), device=self.device, dtype=self.dtype)
        order = [int(x) for x in rep.order[:k]]
        SS = torch.tensor(rep.SS, dtype=torch.float64).view(ms, ms)[order][:, order]
        SY = torch.tensor(rep.SY, dtype=torch.float64).view(ms, ms)[order][:, order]
        M = torch.diag(torch.diag(SY)) + torch.tril(SY, -1) + torch.tril(SY, -1).t() - self._gamma_host * SS
        M = M + self.tol * torch.linalg.norm(M) * torch.eye(k, dtype=torch.float64)
        return M.to(device=self.device, dtype=self.dtype)

    def state_dict(self):
        """delta-independent L-SR1 state for checkpointing."""
        rep = self.ds.read(full=True)
        self.absorb(rep)
        return {"state": self.ds.state.clone(), "S": None if self._Sbuf is None else self._Sbuf.clone(),
                "Y": None if self._Ybuf is None else self._Ybuf.clone(), "n": self.n}

    def load_state_dict(self, sd):
        self.ds.state.copy_(sd["state"])
        if sd["S"] is not None:
            self.ensure_storage(sd["n"])
            self._Sbuf.copy_(sd["S"])
            self._Ybuf.copy_(sd["Y"])
        self.absorb(self.ds.read())
