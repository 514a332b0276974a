"""APTS_D with spatial subdomains for PINNs, drop-in for ``dd4ml.optimizers.apts_pinn``
(apts_pinn.py:7-137): rank r trains on the collocation points in its x-interval (+ overlap)
and weights its local reduction by n_local / N.  Host-only logic on top of APTS_D."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .apts_d import APTS_D


class APTS_PINN(APTS_D):
    __name__ = "APTS_PINN"

    def __init__(self, *args, num_subdomains: int = 1, overlap: float = 0.0, criterion=None, **kwargs):
        super().__init__(*args, nr_models=num_subdomains, criterion=criterion, **kwargs)
        self.num_subdomains = max(1, int(num_subdomains))
        self.domain_low = float(getattr(self.criterion, "low", 0.0))
        self.domain_high = float(getattr(self.criterion, "high", 1.0))
        self.subdomain_bounds = torch.linspace(self.domain_low, self.domain_high, self.num_subdomains + 1)
        self.overlap = max(float(overlap), 0.0)
        self._half_overlap = self.overlap / 2.0
        self._bounds_cache = {}
        self._criterion_attrs = ["current_xt", "current_x"]
        self._sel, self._sel_src, self._sel_ver = None, None, None

    def get_subdomain_bounds(self, idx: int):
        """apts_pinn.py:31-43."""
        if idx not in self._bounds_cache:
            lo, hi = self.subdomain_bounds[idx].item(), self.subdomain_bounds[idx + 1].item()
            if self.overlap > 0.0:
                lo = max(lo - self._half_overlap, self.domain_low)
                hi = min(hi + self._half_overlap, self.domain_high)
            self._bounds_cache[idx] = (lo, hi)
        return self._bounds_cache[idx]

    def get_subdomain_mask(self, x: torch.Tensor, idx: int):
        lo, hi = self.get_subdomain_bounds(idx)
        return (x >= lo) & (x <= hi)

    def _set_criterion_input(self, t):
        for attr in self._criterion_attrs:
            if hasattr(self.criterion, attr):
                setattr(self.criterion, attr, t)

    def step(self, inputs, labels, inputs_d=None, labels_d=None, hNk=None):
        """apts_pinn.py:60-137."""
        idx = dist.get_rank() if dist.is_initialized() else 0
        # boolean-mask indexing (apts_pinn.py:80-90) has a data-dependent shape, i.e. a host sync per
        # call; the row indices are reused only while the SAME tensor object is passed unmodified
        # (a strong reference is kept, so its address cannot be recycled for another point set)
        if self._sel_src is not inputs or self._sel_ver != inputs._version:
            mask = self.get_subdomain_mask(inputs[..., 0].squeeze(), idx)
            self._sel, self._sel_src, self._sel_ver = mask.nonzero().reshape(-1), inputs, inputs._version
        full_in = inputs.detach().requires_grad_(True)
        sub_in = full_in.index_select(0, self._sel)
        sub_lab = labels.index_select(0, self._sel) if labels is not None else None
        weight = int(sub_in.shape[0]) / max(int(full_in.shape[0]), 1)
        self.inputs_d = self.labels_d = None
        self.hNk = hNk
        return self._outer((full_in, labels), (sub_in, sub_lab), weight=weight,
                           on_glob=lambda: self._set_criterion_input(full_in),
                           on_loc=lambda: self._set_criterion_input(sub_in))
