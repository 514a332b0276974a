"""Flat-buffer layout of parameters and gradients.

The reference re-flattens T parameter / gradient tensors with per-tensor copies on every
evaluation (``tr.py:88-131``, ``lssr1_tr.py:205-245``, ``apts_utils.py:17-62``,
``apts_base.py:186-226``).  Here a parameter list is bound ONCE to two flat device
buffers: ``p.data`` and ``p.grad`` become views into them (ordering = the order of the
list, which is what ``parameters_to_vector`` uses), so the flat vectors the kernels
work on are always current and flattening costs 0 bytes.

If a caller replaces ``p.grad`` (``zero_grad(set_to_none=True)``) the views are
restored with one multi-tensor gather kernel; if ``p.data`` is replaced
(``model.double()``, ``load_state_dict`` keeps storage, ``.to(dtype)`` does not) the
buffer is rebuilt and ``version`` is bumped so that owners re-allocate dependent state.
"""
from __future__ import annotations

import ctypes
import weakref

import torch

from . import _lib

_REG: dict[int, tuple[weakref.ref, "FlatBuffer", int]] = {}


class FlatBuffer:
    def __init__(self, params):
        self.params = list(params)
        if not self.params:
            raise ValueError("FlatBuffer needs at least one parameter")
        self.version = 0
        self._bind()

    # ------------------------------------------------------------------ binding
    def _bind(self):
        ps = self.params
        dev, dtype = ps[0].device, ps[0].dtype
        if dev.type != "cuda":
            raise _lib.DD4MLError("dd4ml_b200 has no CPU path: parameters must live on a CUDA device")
        _lib.dt(dtype)
        for p in ps:
            if p.device != dev or p.dtype != dtype:
                raise _lib.DD4MLError("dd4ml_b200: all parameters of one optimizer must share device and dtype")
        self.device, self.dtype = dev, dtype
        self.numels = [p.numel() for p in ps]
        self.offsets = [0]
        for k in self.numels:
            self.offsets.append(self.offsets[-1] + k)
        self.n = self.offsets[-1]
        data = torch.empty(self.n, dtype=dtype, device=dev)
        # two tail elements ride along with the gradient in one all-reduce ([g || loss], APTS)
        self.grad_ext = torch.zeros(self.n + 2, dtype=dtype, device=dev)
        grad = self.grad_ext[: self.n]
        self._gviews = []
        with torch.no_grad():
            for p, off, k in zip(ps, self.offsets, self.numels):
                dv = data[off:off + k].view(p.shape)
                dv.copy_(p.data)
                p.data = dv
                gv = grad[off:off + k].view(p.shape)
                if p.grad is not None:
                    gv.copy_(p.grad)
                if p.requires_grad:
                    p.grad = gv
                self._gviews.append(gv)
                _REG[id(p)] = (weakref.ref(p), self, off)
        self.data, self.grad = data, grad
        self._base = data.data_ptr()
        self.version += 1

    def intact(self) -> bool:
        es = self.data.element_size()
        base = self._base
        for p, off in zip(self.params, self.offsets):
            if p.dtype != self.dtype or p.data_ptr() != base + off * es:
                return False
        return True

    def ensure(self) -> bool:
        """Rebuild after an external storage change.  Returns True when rebuilt."""
        if self.intact():
            return False
        self._bind()
        return True

    # ------------------------------------------------------------------ gradients
    def flat_grad(self) -> torch.Tensor:
        """The flat gradient of the last backward (views kept, or one gather launch)."""
        # nn.Module._apply (.double() / .to(dtype)) swaps `param.data` AND `param.grad.data` in place:
        # the identity test below would still pass while the view objects no longer alias the flat
        # buffer.  A whole-model conversion always moves the first parameter: O(1) check, then rebind.
        p0 = self.params[0]
        if p0.data_ptr() != self._base or p0.dtype != self.dtype:
            self.ensure()
        stale = [i for i, (p, gv) in enumerate(zip(self.params, self._gviews))
                 if p.requires_grad and p.grad is not gv]
        if stale:
            self._regather(stale)
        return self.grad

    def _regather(self, idx):
        lib = _lib.get()
        srcs, offs, lens = [], [], []
        for i in idx:
            p = self.params[i]
            if p.grad is None:
                self._gviews[i].zero_()
            else:
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if g.dtype != self.dtype:
                    g = g.to(self.dtype)
                srcs.append(g)
                offs.append(self.offsets[i])
                lens.append(self.numels[i])
        if srcs:
            n = len(srcs)
            a_src = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
            a_off = (ctypes.c_int64 * n)(*offs)
            a_len = (ctypes.c_int64 * n)(*lens)
            lib.gather(_lib.ptr(self.grad), n, a_src, a_off, a_len, _lib.dt(self.dtype), _lib.stream())
        for i in idx:
            if self.params[i].requires_grad:
                self.params[i].grad = self._gviews[i]

    def rebind_grads(self):
        """Point every ``p.grad`` at its flat view again (no memset): used after a CUDA-graph replay,
        which writes the gradient through the captured views whatever ``p.grad`` currently is."""
        for p, gv in zip(self.params, self._gviews):
            if p.requires_grad and p.grad is not gv:
                p.grad = gv

    def zero_grad(self):
        self.grad.zero_()
        self.rebind_grads()


class FlatSlice:
    """A contiguous run [start, start+n) of a FlatBuffer (e.g. the parameters a rank owns)."""

    def __init__(self, fb: FlatBuffer, start: int, n: int):
        self.fb, self.start, self.n = fb, start, n

    @property
    def version(self):
        return self.fb.version

    @property
    def dtype(self):
        return self.fb.dtype

    @property
    def device(self):
        return self.fb.device

    def ensure(self):
        return self.fb.ensure()

    @property
    def data(self) -> torch.Tensor:
        return self.fb.data[self.start:self.start + self.n]

    def flat_grad(self) -> torch.Tensor:
        return self.fb.flat_grad()[self.start:self.start + self.n]

    def zero_grad(self):
        self.fb.zero_grad()

    def rebind_grads(self):
        self.fb.rebind_grads()


def get_flat(params) -> FlatSlice:
    """FlatSlice for a parameter list: reuses an existing FlatBuffer when the parameters are
    already consecutive views of one, otherwise binds a new buffer."""
    ps = list(params)
    if not ps:
        raise ValueError("optimizer got an empty parameter list")
    hit = [_REG.get(id(p)) for p in ps]
    if all(h is not None and h[0]() is p for h, p in zip(hit, ps)):
        fb = hit[0][1]
        if all(h[1] is fb for h in hit) and fb.intact():
            start = hit[0][2]
            off, ok = start, True
            for h, p in zip(hit, ps):
                ok = ok and h[2] == off
                off += p.numel()
            if ok:
                return FlatSlice(fb, start, off - start)
    fb = FlatBuffer(ps)
    return FlatSlice(fb, 0, fb.n)
