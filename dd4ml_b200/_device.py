"""Device-resident optimizer state block + its pinned host report buffer."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._state import Report

_ERRORS = {
    1: (torch.linalg.LinAlgError, "linalg.cholesky: Psi^T Psi is not positive-definite (solvers/obs.py:65)"),
    2: (torch.linalg.LinAlgError, "linalg.solve: the k x k system is singular (lsr1.py:197 / obs.py:67,177)"),
    3: (ValueError, "OBS input contains NaN or Inf values (solvers/obs.py:48-57)"),
    4: (IndexError, "OBS hard-case branch (solvers/obs.py:100-158) reached; unreachable for gamma > 0"),
    5: (ValueError, "Delta must be non-negative (solvers/obs.py:58)"),
}


class DeviceState:
    """One ``dd4ml_state`` block (csrc/state.h) and reduction workspace on ``device``."""

    def __init__(self, device, gamma0: float = 1.0):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DD4MLError("dd4ml_b200 has no CPU path: optimizer state must live on a CUDA device")
        self.lib = _lib.get()
        self.lay = self.lib.layout
        self.device = device
        with torch.cuda.device(device):
            self.state = torch.empty(self.lib.state_doubles, dtype=torch.float64, device=device)
            self.ws = torch.empty(self.lib.workspace_doubles, dtype=torch.float64, device=device)
            self.lib.state_init(self.state.data_ptr(), self.ws.data_ptr(), float(gamma0), _lib.stream())
        self._host = torch.empty(self.lib.state_doubles, dtype=torch.float64).pin_memory()
        self._np = self._host.numpy()
        self.sp, self.wp = self.state.data_ptr(), self.ws.data_ptr()
        self.syncs = 0

    def set(self, **fields):
        items = [(self.lay.off(k), float(v)) for k, v in fields.items()]
        for i in range(0, len(items), 32):
            chunk = items[i:i + 32]
            n = len(chunk)
            offs = (ctypes.c_int * n)(*[o for o, _ in chunk])
            vals = (ctypes.c_double * n)(*[v for _, v in chunk])
            self.lib.state_set(self.sp, n, offs, vals, _lib.stream())

    def set_array(self, name, values):
        o, c = self.lay.fields[name]
        vals = [float(v) for v in values]
        assert len(vals) <= c
        items = list(zip(range(o, o + len(vals)), vals))
        for i in range(0, len(items), 32):
            chunk = items[i:i + 32]
            n = len(chunk)
            offs = (ctypes.c_int * n)(*[a for a, _ in chunk])
            v = (ctypes.c_double * n)(*[b for _, b in chunk])
            self.lib.state_set(self.sp, n, offs, v, _lib.stream())

    def read(self, full: bool = False) -> Report:
        """Copy the report block (or the whole state) to pinned memory and wait for it.
        This is the only host synchronisation point of the optimizers."""
        n = self.lib.state_doubles if full else self.lay.report_end
        self._host[:n].copy_(self.state[:n], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.syncs += 1
        return Report(self.lay, self._np[:n].tolist())      # a snapshot: later reads do not change it

    @staticmethod
    def raise_for(err: float):
        code = int(err)
        if code:
            exc, msg = _ERRORS.get(code, (RuntimeError, f"dd4ml_b200 device error {code}"))
            raise exc(msg)
