"""Hyper-parameter tables and the sub-problem entry points of
``dd4ml/utility/optimizer_utils.py`` (same names, same values -- the config contract,
SURVEY §8b / row a18)."""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

__all__ = ["get_apts_hparams", "get_lssr1_tr_hparams", "get_lssr1_loc_tr_hparams", "get_tr_hparams",
           "get_loc_tr_hparams", "get_state_dict", "ensure_tensor", "solve_tr_first_order",
           "solve_tr_second_order"]

_TR_GLOB = dict(nu=0.5, inc_factor=1.2, dec_factor=0.9, nu_dec=0.25, nu_inc=0.75)
_TR_LOC = dict(nu=0.45, inc_factor=1.5, dec_factor=0.6, nu_dec=0.3, nu_inc=0.7)
_LSSR1 = dict(gamma=1e-3, mu=0.9, tau_1=0.1, tau_2=0.25, tau_3=0.75, nu_1=0.25, nu_2=0.5, nu_3=0.9,
              nu_4=1.2, c_1=1e-4, c_2=0.9, alpha_max=2.0)


def _radius_scale(config) -> float:
    """optimizer_utils.py:11-15: local radii are divided by the world size for p-norms < inf."""
    ws = dist.get_world_size() if dist.is_initialized() else 1
    return 1.0 / ws if config.norm_type != math.inf else 1.0


def get_state_dict(model):
    return model.module.state_dict() if hasattr(model, "module") else model.state_dict()


def get_apts_hparams(config):                       # optimizer_utils.py:27-39
    return dict(delta=config.delta, min_delta=config.min_delta, nu_dec=0.25, nu_inc=0.75,
                inc_factor=1.2, dec_factor=0.9, max_delta=config.max_delta)


def _lssr1(config, scale, second_order, dogleg, sync):
    hp = dict(delta=config.delta * scale, min_delta=config.min_delta * scale,
              max_delta=config.max_delta * scale, second_order=second_order, dogleg=dogleg,
              mem_length=config.mem_length, max_wolfe_iters=config.max_wolfe_iters,
              max_zoom_iters=config.max_zoom_iters, tol=config.tol, norm_type=config.norm_type,
              sync=sync, paper_tr_update=config.paper_tr_update)
    hp.update(_LSSR1)
    return hp


def get_lssr1_tr_hparams(config):                   # optimizer_utils.py:42-72
    return _lssr1(config, 1.0, config.glob_second_order, config.glob_dogleg, True)


def get_lssr1_loc_tr_hparams(config):               # optimizer_utils.py:75-106
    return _lssr1(config, _radius_scale(config), config.loc_second_order, config.loc_dogleg, False)


def _tr(config, scale, table, second_order, dogleg):
    hp = dict(delta=config.delta * scale, max_delta=config.max_delta * scale,
              min_delta=config.min_delta * scale, mem_length=5, norm_type=config.norm_type,
              second_order=second_order, dogleg=dogleg, tol=config.tol)
    hp.update(table)
    return hp


def get_tr_hparams(config):                         # optimizer_utils.py:109-128 (mem_length fixed to 5, Q3)
    return _tr(config, 1.0, _TR_GLOB, config.glob_second_order, config.glob_dogleg)


def get_loc_tr_hparams(config):                     # optimizer_utils.py:131-151
    return _tr(config, _radius_scale(config), _TR_LOC, config.loc_second_order, config.loc_dogleg)


def get_loc_tradam_hparams(config):                 # optimizer_utils.py:154-164
    return {"lr": config.learning_rate, "betas": (0.9, 0.999), "eps": 1e-8, "norm_type": config.norm_type}


def ensure_tensor(d, device):
    if torch.is_tensor(d):
        return d.to(device)
    if isinstance(d, (int, float)):
        return torch.scalar_tensor(d, device=device)
    raise TypeError(f"Unexpected type for d: {type(d)}. Expected int, float, or torch.Tensor.")


def _scratch_state(gradient, norm_inf):
    from .._device import DeviceState
    ds = DeviceState(gradient.device)
    ds.set(norm_inf=1.0 if norm_inf else 0.0)
    return ds


def solve_tr_first_order(gradient, grad_norm, trust_radius, tol):
    """optimizer_utils.py:197-215 on the device: step = -g * delta/||g||, pred = delta*||g||.
    ``grad_norm`` is recomputed by the kernel (2-norm, or inf-norm when it equals max|g|)."""
    from .. import _lib
    g = gradient.detach().contiguous()
    gn = float(grad_norm)
    norm_inf = abs(gn - float(g.abs().max())) <= 1e-6 * max(gn, 1e-30) and g.numel() > 1
    ds = _scratch_state(g, norm_inf)
    ds.set(delta=float(trust_radius), tol=float(tol), second_order=0.0)
    p = torch.empty_like(g)
    n, vt = g.numel(), _lib.dt(g.dtype)
    ds.lib.tr_propose(ds.sp, ds.wp, _lib.ptr(g), None, None, n, n, vt, _lib.F32, 0, 0, _lib.stream())
    ds.lib.tr_apply(ds.sp, ds.wp, _lib.ptr(g), None, None, _lib.ptr(p), None, n, n, vt, _lib.F32, vt, 0, _lib.stream())
    rep = ds.read()
    if rep.converged:
        return torch.zeros_like(g), 0.0
    return p, torch.tensor(rep.pred, dtype=g.dtype, device=g.device)


def solve_tr_second_order(gradient, grad_norm, trust_radius, lsr1_hessian, obs_solver, tol, dogleg=False):
    """optimizer_utils.py:218-307 on the device: OBS step, Cauchy point, optional dogleg and
    the predicted reduction -- one reduction pass, the k-space solve, one output pass."""
    from .. import _lib
    from .._device import DeviceState
    g = gradient.detach().contiguous()
    h = lsr1_hessian
    ds = h.ds
    n, vt = g.numel(), _lib.dt(g.dtype)
    h.ensure_storage(n)
    gn = float(grad_norm)
    norm_inf = abs(gn - float(g.abs().max())) <= 1e-6 * max(gn, 1e-30) and n > 1
    ds.set(delta=float(trust_radius), tol=float(tol), dogleg=1.0 if dogleg else 0.0, second_order=1.0,
           norm_inf=1.0 if norm_inf else 0.0)
    p = torch.empty_like(g)
    ds.lib.tr_propose(ds.sp, ds.wp, _lib.ptr(g), h.S_ptr, h.Y_ptr, n, h.ld, vt, _lib.dt(h.dtype), 0, h.memory_length, _lib.stream())
    ds.lib.tr_apply(ds.sp, ds.wp, _lib.ptr(g), h.S_ptr, h.Y_ptr, _lib.ptr(p), None, n, h.ld, vt,
                    _lib.dt(h.dtype), vt, h.memory_length, _lib.stream())
    rep = ds.read()
    DeviceState.raise_for(rep.err)
    h.absorb(rep)
    if obs_solver is not None:
        obs_solver.last = dict(case=int(rep.case_id), sigma=rep.sigma, sigma0=rep.sigma0, tau=rep.tau,
                               newton_iters=int(rep.newton_iters), newton_degenerate=int(rep.newton_deg))
    if rep.converged:
        return torch.zeros_like(g), 0.0
    return p, torch.tensor(rep.pred, dtype=g.dtype, device=g.device)
