"""Oracle: Orthonormal-Basis SR1 trust-region sub-problem solver
(reference ``solvers/obs.py:46-254``).  TEST INFRASTRUCTURE ONLY.

The step is exactly what the reference code computes, including
  * the SMW formula *without* the paper's 1/tau on the Psi-term (obs.py:175-182),
  * the Newton "degenerate" return (-1/delta, 1/tol) (obs.py:242-245),
  * the 200-iteration cap (obs.py:211).
"""
from __future__ import annotations

import torch

OBS_TOL = 1e-6  # obs.py:29 (fixed, independent of the optimizer's tol)


def _phi_bar_f(sigma, Lam, a, delta, tol=OBS_TOL):
    """obs.py:184-208."""
    D = Lam + sigma
    if (a.abs() < tol).any() or (D.abs() < tol).any():
        return -1.0 / delta
    pn2 = 0
    for i in range(a.shape[0]):
        if a[i].abs() > tol and D[i].abs() > tol:
            pn2 = pn2 + (a[i] / D[i]) ** 2
    return 1.0 / torch.sqrt(pn2) - 1.0 / delta


def _phi_bar_fg(sigma, Lam, a, delta, tol=OBS_TOL):
    """obs.py:228-254.  Returns (phi, dphi, degenerate_flag)."""
    D = Lam + sigma
    if (a.abs() < tol).any() or (D.abs() < tol).any():
        return torch.tensor(-1) / delta, torch.tensor(1) / tol, True
    p = a / D
    nrm = torch.norm(p)
    phi = 1 / nrm - 1 / delta
    dphi = torch.sum((a ** 2) / (D ** 3)) / (nrm ** 3)
    return phi, dphi, False


def _newton(x0, Lam, a, delta, tol=OBS_TOL, max_iter=200):
    """obs.py:210-226.  Returns (sigma, iterations, degenerate_evals)."""
    x = x0
    f, g, deg = _phi_bar_fg(x, Lam, a, delta, tol)
    k, ndeg = 0, int(deg)
    while torch.abs(f) > tol and k < max_iter:
        x = x - f / g
        f, g, deg = _phi_bar_fg(x, Lam, a, delta, tol)
        ndeg += int(deg)
        k += 1
    return x, k, ndeg


def canonical_eigvec_signs(U: torch.Tensor) -> torch.Tensor:
    """Flip every column so its largest-|.| entry (lowest index on ties) is > 0."""
    idx = U.abs().argmax(dim=0)
    sgn = torch.sign(U[idx, torch.arange(U.shape[1], device=U.device)])
    sgn[sgn == 0] = 1
    return U * sgn


def obs_solve(g, delta, gamma, Psi, Minv, *, eig_sign="lapack", info=None, tol=OBS_TOL):
    """p* of  min g^T p + 1/2 p^T (gamma I + Psi M Psi^T) p, ||p|| <= delta,
    as the reference computes it (obs.py:46-173).

    ``delta`` and ``gamma`` are 0-d tensors.  ``info`` (dict) receives
    diagnostics: case ("A"|"B"|"C"), sigma, newton_iters, Lambda, a.
    """
    if info is None:
        info = {}
    for name, t in (("g", g), ("delta", delta), ("gamma", gamma), ("Psi", Psi), ("Minv", Minv)):
        if torch.isnan(t).any() or torch.isinf(t).any():      # :48-57
            raise ValueError(f"{name} contains NaN or Inf values")
    if delta < 0:
        raise ValueError("Delta must be non-negative")

    PsiPsi = Psi.t() @ Psi                                     # :61
    R = torch.linalg.cholesky(PsiPsi, upper=True)             # :65
    MR = torch.linalg.solve(Minv, R.t())
    RMR = R @ MR
    RMR = (RMR + RMR.t()) / 2.0
    D, U = torch.linalg.eigh(RMR)                              # :71
    order = torch.argsort(D)
    D, U = D[order], U[:, order]
    if eig_sign == "canonical":
        U = canonical_eigvec_signs(U)
    k = D.shape[0]
    Lam = torch.cat((D + gamma, gamma.view(1).to(D.device)))
    Lam[Lam.abs() < tol] = 0                                   # :79
    lam_min = torch.min(Lam[0], gamma)                         # :80
    RinvU = torch.linalg.solve(R, U)
    gc = g.to(Psi.dtype)
    PsiTg = Psi.t() @ gc
    g_par = RinvU.t() @ PsiTg
    a_last = torch.sqrt(torch.clamp(g.dot(g) - g_par.dot(g_par), min=0.0))
    a = torch.cat((g_par, a_last.view(-1)))
    info.update(Lambda=Lam.clone(), a=a.clone(), lam_min=float(lam_min))

    def smw(tau):                                              # :175-182
        W = tau * Minv + PsiPsi
        return (-1.0 / tau) * gc + Psi @ torch.linalg.solve(W, PsiTg)

    if lam_min > 0 and torch.norm(a / Lam) <= delta:           # :97  case A
        info.update(case="A", sigma=0.0, newton_iters=0)
        return smw(gamma)
    if lam_min <= 0 and _phi_bar_f(-lam_min, Lam, a, delta, tol) >= 0:   # :100 case B (hard case)
        sig = -lam_min
        v = torch.zeros(k + 1, dtype=a.dtype, device=a.device)
        nz = (Lam + sig).abs() > tol
        v[nz] = a[nz] / (Lam[nz] + sig)
        P_par = Psi @ RinvU
        if torch.abs(gamma + sig) < tol:
            p = -1.0 * (P_par @ v[:k])
        else:
            p = (-1.0 * (P_par @ v[:k])
                 + 1.0 / (gamma + sig) * (Psi @ torch.linalg.solve(PsiPsi, PsiTg))
                 - g / (gamma + sig))
        if lam_min < 0:
            alpha = torch.sqrt(torch.clamp(delta ** 2 - p.dot(p), min=0.0))
            if torch.abs(lam_min - Lam[0]) < tol:              # :127
                p = p + (1.0 / torch.norm(P_par[:, 0])) * alpha * P_par[:, 0]
            else:
                # obs.py:132-154 indexes a 1-D tensor with .transpose(0, 1) and would
                # raise; it is unreachable because gamma >= tol > 0.
                raise IndexError("OBS hard-case fallback branch (unreachable in the reference)")
        info.update(case="B", sigma=float(sig), newton_iters=0)
        return p
    # case C: Newton on the secular equation                 # :159-172
    if lam_min > 0:
        x0 = 0
    else:
        sig_hat = torch.max(a / delta - Lam)
        x0 = sig_hat if sig_hat > -lam_min else -lam_min
    sig, iters, ndeg = _newton(x0, Lam, a, delta, tol)
    sig = torch.as_tensor(sig, dtype=a.dtype)
    if torch.isnan(sig) or torch.isinf(sig):
        sig, iters, ndeg = _newton(0, Lam, a, delta, tol)
        sig = torch.as_tensor(sig, dtype=a.dtype)
    info.update(case="C", sigma=float(sig), newton_iters=iters, newton_degenerate=ndeg,
                x0=float(x0))
    return smw(gamma + sig)
