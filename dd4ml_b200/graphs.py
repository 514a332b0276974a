"""CUDA-graph replay of a loss/gradient evaluation.

The closures the APTS optimizers call (``apts_base.py:245-318``) are a fixed sequence of
PyTorch kernels for a fixed (model, criterion, batch shape): zero the flat gradient buffer,
forward, loss, backward into the flat gradient views.  On a B200 that sequence is CPU
launch-bound for the BASELINE models (tens of microseconds of GPU work behind ~1 ms of
Python/dispatcher time), so it is captured once into a CUDA graph and replayed -- the
model's forward/backward kernels themselves are untouched.

Inputs are captured by address.  When the caller keeps passing the same tensors the graph
reads them in place; a small ring of batch buffers (``HostBatchPrefetcher``: two) gets one graph
per buffer, all in one memory pool; once more than ``MAX_ADDRESSES`` distinct addresses have been
seen the evaluator switches to a private static copy (one ``copy_`` per call).  Anything that cannot be captured (data
dependent Python control flow in the model, unsupported ops) falls back to the eager
PyTorch path for that shape -- this only concerns the PyTorch closure, never the
dd4ml_b200 kernels.
"""
from __future__ import annotations

import os

import torch

ENABLED = os.environ.get("DD4ML_B200_CUDA_GRAPHS", "1") != "0"
MAX_ADDRESSES = 3     # in-place graphs per batch shape before falling back to a private input copy
MAX_SHAPES = 4        # batch shapes kept (the Trainer's batch-size growth loop keeps producing new ones)


class _Entry:
    __slots__ = ("graph", "static_in", "static_lab", "loss", "private", "calls")


class GraphedEval:
    PROF = None       # bench.py: list of (start_event, end_event) per evaluation when profiling

    def __call__(self, inputs, labels):
        if GraphedEval.PROF is None:
            return self._call(inputs, labels)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = self._call(inputs, labels)
        e1.record()
        GraphedEval.PROF.append((e0, e1))
        return out

    def __init__(self, model, criterion, flat, long_labels=True, warmup=3):
        self.model, self.criterion, self.flat = model, criterion, flat
        self.long_labels = long_labels
        self.warmup = warmup
        self.cache: dict = {}
        self.failed: set = set()
        self.replays = 0
        self.captures = 0
        self._hot = None      # (inputs, labels, flat version, training, in ptr, lab ptr, entry) of the last replay

    # the eager evaluation (also what gets captured)
    def _eval(self, inputs, labels):
        self.flat.zero_grad()
        lab = labels.long() if (self.long_labels and torch.is_tensor(labels)) else labels
        loss = self.criterion(self.model(inputs), lab)
        # inputs that carry autograd history (APTS_PINN: sub_in = full_in[mask]) are evaluated
        # several times per step; like the reference (apts_base.py:258) keep their graph alive
        loss.backward(retain_graph=bool(torch.is_tensor(inputs) and inputs.requires_grad))
        return loss.detach()

    def _key(self, inputs, labels):
        # everything a captured graph bakes in besides the tensor addresses: shapes / dtypes / strides
        # of the batch, the flat-buffer generation (a dtype or storage change re-binds the parameters:
        # an old graph would read the orphaned storage) and train/eval mode (dropout, BatchNorm)
        return (tuple(inputs.shape), inputs.dtype, tuple(inputs.stride()), tuple(labels.shape), labels.dtype,
                tuple(labels.stride()), self.flat.version, self.model.training)

    def _capture(self, inputs, labels, private, pool=None):
        e = _Entry()
        e.private = private
        e.static_in = inputs.clone() if private else inputs
        e.static_lab = labels.clone() if private else labels
        e.calls = 0
        # warm-up evaluations must not leave side effects (BatchNorm running statistics ...)
        saved = [b.detach().clone() for b in self.model.buffers()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self.warmup):
                self._eval(e.static_in, e.static_lab)
        torch.cuda.current_stream().wait_stream(s)
        with torch.no_grad():
            for b, v in zip(self.model.buffers(), saved):
                b.copy_(v)
        e.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e.graph, pool=pool):       # graphs of one shape never run concurrently
            e.loss = self._eval(e.static_in, e.static_lab)
        self.captures += 1
        return e

    def _call(self, inputs, labels):
        # hot path (every evaluation of an outer iteration passes the same batch objects): the replay is the
        # first device work after a host read, so nothing but these identity checks precedes it
        hot = self._hot
        if (hot is not None and hot[0] is inputs and hot[1] is labels and hot[2] == self.flat.version
                and hot[3] == self.model.training and hot[4] == inputs.data_ptr() and hot[5] == labels.data_ptr()
                and not inputs.requires_grad):
            e = hot[6]
            e.graph.replay()
            self.flat.rebind_grads()
            self.replays += 1
            return e.loss.clone()
        if (not ENABLED or not torch.is_tensor(inputs) or not torch.is_tensor(labels) or inputs.requires_grad
                or not inputs.is_cuda):
            return self._eval(inputs, labels)
        key = self._key(inputs, labels)
        if self.cache and next(iter(self.cache))[-2] != key[-2]:
            self.cache.clear()                   # parameters were re-bound: drop every stale graph
        if key in self.failed:
            return self._eval(inputs, labels)
        ents = self.cache.get(key)
        if ents is None:
            while len(self.cache) >= MAX_SHAPES:         # oldest shape first (dicts keep insertion order)
                old = next(iter(self.cache))
                if self._hot is not None and any(self._hot[6] is x for x in self.cache[old]):
                    self._hot = None
                del self.cache[old]                      # frees the graphs and, with them, their memory pool
            ents = self.cache[key] = []
        ip, lp = inputs.data_ptr(), labels.data_ptr()
        e = next((x for x in ents if not x.private and x.static_in.data_ptr() == ip and x.static_lab.data_ptr() == lp),
                 None)
        if e is None:
            e = next((x for x in ents if x.private), None)     # (only exists once the address budget is spent)
        try:
            if e is None:
                pool = ents[0].graph.pool() if ents else None
                e = self._capture(inputs, labels, private=len(ents) >= MAX_ADDRESSES, pool=pool)
                ents.append(e)
        except Exception:  # noqa: BLE001  (capture unsupported for this model: eager PyTorch closure)
            self.failed.add(key)
            torch.cuda.synchronize()
            return self._eval(inputs, labels)
        if e.private:
            e.static_in.copy_(inputs)
            e.static_lab.copy_(labels)
            self._hot = None
        else:
            self._hot = (inputs, labels, key[-2], key[-1], ip, lp, e)
        e.graph.replay()
        # the replay wrote the gradient into the flat views; a caller may have set p.grad to None
        # (zero_grad(set_to_none=True)) since the capture -- re-establish the views so that
        # flat_grad() does not take the fresh gradient for a missing one
        self.flat.rebind_grads()
        self.replays += 1
        return e.loss.clone()
