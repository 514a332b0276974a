"""Model cloning / parameter-ownership helpers (``dd4ml/utility/apts_utils.py``)."""
from __future__ import annotations

import copy

import torch
import torch.distributed as dist
import torch.nn as nn

from .optimizer_utils import get_state_dict

__all__ = ["clone_model", "broadcast_shuffle", "split_indices", "mark_trainable", "flatten_params",
           "restore_params", "trainable_params_to_vector", "trainable_grads_to_vector"]


def clone_model(model):
    """apts_utils.py:71-80: a structurally identical copy with the same weights.  Models that
    carry a ``config`` are rebuilt from it like the reference does; anything else is deep-copied."""
    base = model.module if hasattr(model, "module") else model
    ref = next(model.parameters())
    if hasattr(base, "config"):
        cfg = copy.deepcopy(base.config)
        if hasattr(cfg, "model_type"):
            cfg.model_type = None
        new = type(base)(cfg).to(device=ref.device, dtype=ref.dtype)
        new.load_state_dict(get_state_dict(base))
        return new
    new = copy.deepcopy(base)
    # deep copies of flat-buffer views share no storage with the original
    return new.to(device=ref.device, dtype=ref.dtype)


def broadcast_shuffle(num_layers: int, rank: int, world_size: int, device=None) -> torch.Tensor:
    """apts_utils.py:83-90: rank 0 draws the permutation of parameter tensors, everyone gets it."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    if rank == 0:
        perm = torch.randperm(num_layers, dtype=torch.long, device=device)
    else:
        perm = torch.empty(num_layers, dtype=torch.long, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.broadcast(perm, src=0)
    return perm


def split_indices(perm, rank: int, world_size: int):
    """apts_utils.py:93-98: contiguous chunk of the permutation, sizes differ by at most one."""
    q, r = divmod(len(perm), world_size)
    start = rank * q + min(rank, r)
    return perm[start:start + q + (1 if rank < r else 0)]


def mark_trainable(model: nn.Module):
    """apts_utils.py:101-124: each rank trains a random disjoint subset of parameter tensors."""
    if dist.is_available() and dist.is_initialized():
        rank, world_size = dist.get_rank(), dist.get_world_size()
    else:
        rank, world_size = 0, 1
    params = list(model.parameters())
    if not params:
        return model
    if world_size > len(params):
        raise ValueError("World size cannot be greater than the number of layers in the model.")
    perm = broadcast_shuffle(len(params), rank, world_size, device=params[0].device)
    loc = split_indices(perm, rank, world_size)
    model._trainable_indices = loc
    owned = set(int(i) for i in loc)
    for i, p in enumerate(params):
        p.requires_grad = i in owned
    return model


def flatten_params(model, out=None):
    """apts_utils.py:17-58 (compatibility helper; the optimizers use flat-buffer views)."""
    v = torch.nn.utils.parameters_to_vector(model.parameters()).detach()
    if out is not None:
        out.copy_(v)
        return out.clone()
    return v


def restore_params(model, flat_params):
    torch.nn.utils.vector_to_parameters(flat_params, model.parameters())


def trainable_params_to_vector(model: nn.Module) -> torch.Tensor:
    ps = [p for p in model.parameters() if p.requires_grad]
    if not ps:
        return torch.tensor([], dtype=torch.float32)
    return torch.cat([p.detach().reshape(-1) for p in ps]).clone()


def trainable_grads_to_vector(model: nn.Module) -> torch.Tensor:
    ps = [p for p in model.parameters() if p.requires_grad]
    if not ps:
        return torch.tensor([], dtype=torch.float32)
    return torch.cat([p.grad.reshape(-1) if p.grad is not None else torch.zeros_like(p).reshape(-1)
                      for p in ps]).detach()
