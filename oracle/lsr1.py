"""Oracle: compact limited-memory SR1 (reference ``optimizers/lsr1.py``).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

B = gamma*I + Psi * M * Psi^T,  Psi = Y - gamma*S,
M^{-1} = diag(S^T Y) + L + L^T - gamma * S^T S  (+ tol*||M^{-1}||_F * I),
L = strict lower triangle of S^T Y.
"""
from __future__ import annotations

import torch


class LSR1Memory:
    """State of the L-SR1 approximation (lsr1.py:26-51).

    Pairs are kept as a python list of flat vectors, oldest first; the
    reference's (n x m) workspace with a column shift (lsr1.py:112-131) is the
    same sequence.
    """

    def __init__(self, gamma=1.0, memory_length=10, tol=1e-6, dtype=torch.float32):
        self.m = int(memory_length)
        self.tol = float(tol)
        self.dtype = dtype
        self.gamma = torch.tensor(float(gamma), dtype=dtype)  # lsr1.py:38
        self.S: list[torch.Tensor] = []
        self.Y: list[torch.Tensor] = []
        self.Psi = None   # (n, k) after precompute
        self.Minv = None  # (k, k) after precompute

    def __len__(self):
        return len(self.S)

    # -- lsr1.py:181-198 ---------------------------------------------------
    def B(self, v: torch.Tensor) -> torch.Tensor:
        v = v.to(self.dtype)
        if self.Psi is None or self.Psi.numel() == 0:
            return self.gamma * v
        rhs = self.Psi.t() @ v
        x = torch.linalg.solve(self.Minv, rhs)
        return self.gamma * v + self.Psi @ x

    # -- lsr1.py:53-144 ----------------------------------------------------
    def update(self, s: torch.Tensor, y: torch.Tensor) -> bool:
        """Try to append (s, y).  Returns True when the pair was stored."""
        s = s.flatten().to(self.dtype)
        y = y.flatten().to(self.dtype)
        s_norm, y_norm = s.norm(), y.norm()
        if s_norm <= self.tol or y_norm <= self.tol:          # :72
            return False
        curvature = y.dot(s)
        if abs(curvature) <= self.tol * (s_norm * y_norm):     # :78
            return False
        # B(s) uses Psi/Minv of the last precompute and the current gamma (:83)
        u = y - self.B(s)
        denom = u.dot(s)
        if abs(denom) <= self.tol * s.norm() * u.norm():      # :86
            return False
        gamma_cand = (y.dot(y) / curvature).clamp_min(self.tol)  # :90
        psi_cand = y - gamma_cand * s
        if psi_cand.norm() <= self.tol * y.norm():            # :93
            return False
        if len(self.S) >= self.m:                             # :122-140 (drop oldest)
            self.S.pop(0)
            self.Y.pop(0)
        self.S.append(s)
        self.Y.append(y)
        self.gamma = gamma_cand                               # :144
        return True

    # -- lsr1.py:146-179 ---------------------------------------------------
    def precompute(self) -> None:
        k = len(self.S)
        if k == 0:
            self.Psi = torch.zeros((0, 0), dtype=self.dtype)
            self.Minv = torch.zeros((0, 0), dtype=self.dtype)
            return
        # (n, m) workspace with the k live columns as a strided view -- the same memory
        # layout the reference hands to MKL (lsr1.py:112-120,158-160), so that the GEMM
        # rounding is reproduced as well.
        n = self.S[0].shape[0]
        dev = self.S[0].device                     # (cpu everywhere except bench.py --impl aten)
        Sm = torch.zeros(n, self.m, dtype=self.dtype, device=dev)
        Ym = torch.zeros(n, self.m, dtype=self.dtype, device=dev)
        for j in range(k):
            Sm[:, j] = self.S[j]
            Ym[:, j] = self.Y[j]
        S, Y = Sm[:, :k], Ym[:, :k]
        self.Psi = Y - self.gamma * S
        SY = S.t() @ Y
        Minv = (torch.diag(torch.diag(SY)) + torch.tril(SY, -1)
                + torch.tril(SY, -1).t() - self.gamma * (S.t() @ S))
        Minv = Minv + (self.tol * torch.norm(Minv, p="fro")) * torch.eye(k, dtype=self.dtype, device=dev)
        self.Minv = Minv
