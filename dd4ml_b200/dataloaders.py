"""Overlap / micro-batch samplers and a GPU-resident batch feeder.

The three sampler classes keep the names, constructor arguments, error behaviour and -- index for
index -- the batch sequences of the reference (dd4ml/dataloaders.py:18-229; used at
trainer.py:356-382), so they can be handed to a ``torch.utils.data.DataLoader`` exactly like the
reference's.  They are written as closed-form index arithmetic over the cached base order instead
of per-element Python loops, which is what lets ``DeviceBatchFeeder`` precompute a whole epoch of
batches as ONE device index tensor: the data set stays in HBM and every batch is one launch of the
row-gather kernel (``dd4ml_gather_rows``, csrc/k_misc.cu) -- no host->device copy and no
Python-side collation per step, so the outer iteration is not input-bound.
"""
from __future__ import annotations

import math
from typing import Iterator, List, Sequence, Union

import numpy as np
import torch

from . import _lib


def _overlap_size(overlap, batch_size: int) -> int:
    """dataloaders.py:37-42: a float in (0, 1) is a fraction of the batch, an int >= 1 a count."""
    if isinstance(overlap, float) and 0.0 < overlap < 1.0:
        return int(round(overlap * batch_size))
    if isinstance(overlap, int) and overlap >= 1:
        return min(overlap, batch_size - 1)
    return 0


class OverlapBatchSampler:
    """Successive mini-batches share ``overlap`` examples without changing the number of batches:
    batch b = base[b*bs : (b+1)*bs] followed by the next ``overlap_sz`` positions (circular).  With
    ``drop_last`` the last batch takes its overlap from the start of the order; without it a final
    partial batch is yielded as is, without overlap (dataloaders.py:57-96)."""

    def __init__(self, base_sampler, batch_size: int, overlap: Union[float, int] = 0.0, drop_last: bool = False):
        self.base_sampler = base_sampler
        self.batch_size = int(batch_size)
        self.drop_last = drop_last
        self.overlap_sz = _overlap_size(overlap, self.batch_size)
        if self.overlap_sz >= self.batch_size:
            raise ValueError("`overlap` must be smaller than `batch_size`")
        self._cached_indices = None      # the base order is drawn once and reused by every epoch

    def _get_indices(self):
        if self._cached_indices is None:
            self._cached_indices = list(self.base_sampler)
        return self._cached_indices

    def batch_positions(self) -> List[np.ndarray]:
        """Positions into the cached base order of every batch (what ``__iter__`` maps through it)."""
        n = len(self._get_indices())
        if n == 0:
            return []
        bs, ov = self.batch_size, self.overlap_sz
        nb = n // bs if self.drop_last else math.ceil(n / bs)
        out = []
        for b in range(nb):
            start, end = b * bs, min((b + 1) * bs, n)
            uniq = np.arange(start, end, dtype=np.int64)
            if end - start < bs or ov == 0:       # final partial batch (only without drop_last) / no overlap
                out.append(uniq)
                continue
            o0 = 0 if (self.drop_last and b == nb - 1) else end % n
            out.append(np.concatenate([uniq, (o0 + np.arange(ov, dtype=np.int64)) % n]))
        return out

    def __iter__(self) -> Iterator[List[int]]:
        idx = np.asarray(self._get_indices(), dtype=np.int64)
        for pos in self.batch_positions():
            yield idx[pos].tolist()

    def __len__(self):
        n = len(self.base_sampler) if self._cached_indices is None else len(self._cached_indices)
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)


class MicroBatchOverlapSampler:
    """Splits every mini-batch of an ``OverlapBatchSampler`` into ``num_subdomains`` micro-batches:
    the unique part and the overlap part are each dealt round-robin (dataloaders.py:148-168)."""

    def __init__(self, overlap_sampler: OverlapBatchSampler, num_subdomains: int,
                 allow_empty_microbatches: bool = False):
        if not isinstance(overlap_sampler, OverlapBatchSampler):
            raise TypeError("overlap_sampler must be an instance of OverlapBatchSampler")
        if num_subdomains <= 0:
            raise ValueError("num_subdomains must be positive")
        if overlap_sampler.overlap_sz == 0:
            raise ValueError("OverlapBatchSampler must have overlap > 0 for micro-batch overlap")
        total = overlap_sampler.batch_size + overlap_sampler.overlap_sz
        if num_subdomains > total and not allow_empty_microbatches:
            raise ValueError(f"num_subdomains ({num_subdomains}) > total mini-batch size ({total}).")
        self.overlap_sampler = overlap_sampler
        self.num_subdomains = num_subdomains
        self.allow_empty_microbatches = allow_empty_microbatches

    def _split_mini_batch(self, mini_batch: Sequence[int]) -> List[List[int]]:
        mb = list(mini_batch)
        nu = min(len(mb), self.overlap_sampler.batch_size)
        uniq, over = mb[:nu], mb[nu:]
        N = self.num_subdomains
        parts = [uniq[j::N] + over[j::N] for j in range(N)]
        if not self.allow_empty_microbatches:
            empties = sum(1 for p in parts if not p)
            if empties:
                raise RuntimeError(f"{empties} empty micro-batches generated.")
        return parts

    def __iter__(self) -> Iterator[List[List[int]]]:
        for mb in self.overlap_sampler:
            yield self._split_mini_batch(mb)

    def __len__(self):
        return len(self.overlap_sampler)

    def get_overlap_info(self) -> dict:
        """dataloaders.py:173-204."""
        info = {"mini_batches": len(self.overlap_sampler), "subdomains": self.num_subdomains,
                "overlap_per_minibatch": self.overlap_sampler.overlap_sz,
                "allow_empty_microbatches": self.allow_empty_microbatches}
        mbs = list(self.overlap_sampler)
        if len(mbs) >= 2:
            m0, m1, ml = (self._split_mini_batch(mbs[i]) for i in (0, 1, -1))
            info.update({
                "actual_microbatch_overlaps": [len(set(a) & set(b)) for a, b in zip(m0, m1)],
                "first_microbatch_sizes": [len(m) for m in m0],
                "second_microbatch_sizes": [len(m) for m in m1],
                "last_microbatch_sizes": [len(m) for m in ml],
            })
        else:
            info["note"] = "Need ≥2 mini-batches to calc overlap."
        return info


class MicroBatchFlattenSampler:
    """Flattens the micro-batches into individual batches for a DataLoader (dataloaders.py:207-224)."""

    def __init__(self, micro_batch_sampler: MicroBatchOverlapSampler):
        self.micro_batch_sampler = micro_batch_sampler

    def __iter__(self) -> Iterator[List[int]]:
        for parts in self.micro_batch_sampler:
            for p in parts:
                if p:
                    yield p

    def __len__(self):
        return len(self.micro_batch_sampler) * self.micro_batch_sampler.num_subdomains


class DeviceBatchFeeder:
    """Batches of a data set that lives in HBM, selected by any of the batch samplers above.

    ``tensors``: the data set as tensors with a common first dimension (inputs, labels, ...),
    moved to ``device`` once.  The sampler's batches of one epoch are flattened into one device
    index vector at construction (the reference's samplers cache their base order, so every epoch
    repeats them); iterating yields tuples of device tensors, each filled by one row-gather launch.
    """

    def __init__(self, tensors: Sequence[torch.Tensor], batch_sampler, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DD4MLError("dd4ml_b200 has no CPU path: DeviceBatchFeeder needs a CUDA device")
        self.lib = _lib.get()
        self.data = [t.to(self.device).contiguous() for t in tensors]
        n = self.data[0].shape[0]
        if any(t.shape[0] != n for t in self.data):
            raise ValueError("all tensors must have the same first dimension")
        batches = [np.asarray(b, dtype=np.int64) for b in batch_sampler]
        for b in batches:
            if b.size and (b.min() < 0 or b.max() >= n):
                raise IndexError("batch sampler produced an index outside the data set")
        self.sizes = [int(b.size) for b in batches]
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        flat = np.concatenate(batches) if batches else np.zeros(0, dtype=np.int64)
        self.index = torch.from_numpy(flat).to(self.device)
        self._row_bytes = [int(t[0].numel()) * t.element_size() if n else 0 for t in self.data]

    def __len__(self):
        return len(self.sizes)

    def batch(self, i: int, out: Sequence[torch.Tensor] | None = None):
        """Gather batch ``i``; ``out`` (optional) are preallocated destination tensors."""
        rows, off = self.sizes[i], int(self.offsets[i])
        res = []
        for j, (t, rb) in enumerate(zip(self.data, self._row_bytes)):
            dst = out[j] if out is not None else torch.empty((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
            if dst.shape[0] != rows or not dst.is_contiguous():
                raise ValueError("destination tensor does not match the batch")
            if rows and rb:
                self.lib.gather_rows(_lib.ptr(dst), _lib.ptr(t), self.index.data_ptr() + 8 * off, rows, rb, t.shape[0],
                                     _lib.stream())
            res.append(dst)
        return tuple(res)

    def __iter__(self):
        for i in range(len(self)):
            yield self.batch(i)


class HostBatchPrefetcher:
    """Host-resident (pinned) batches -> device, with the copy of batch i+1 overlapped with the
    optimizer step on batch i: a side stream fills a ring of ``depth`` device buffers, events order
    it against the compute stream in both directions.  ``source`` is an iterable of tuples of
    pinned host tensors of fixed shapes (e.g. a ``DataLoader(..., pin_memory=True)``)."""

    def __init__(self, source, device="cuda", depth: int = 2):
        self.source = source
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DD4MLError("dd4ml_b200 has no CPU path: HostBatchPrefetcher needs a CUDA device")
        self.depth = max(2, int(depth))

    def __iter__(self):
        copy_stream = torch.cuda.Stream(self.device)
        ring, ready, freed = [], [], []
        it = iter(self.source)

        def issue(slot, host):
            if slot == len(ring):
                ring.append(tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host))
                ready.append(torch.cuda.Event())
                freed.append(None)
            with torch.cuda.stream(copy_stream):
                if freed[slot] is not None:
                    copy_stream.wait_event(freed[slot])      # the consumer of this buffer has finished
                for d, h in zip(ring[slot], host):
                    d.copy_(h, non_blocking=True)
                ready[slot].record(copy_stream)

        pending = []                                         # slots in flight, oldest first
        nxt = 0
        for _ in range(self.depth - 1):
            host = next(it, None)
            if host is None:
                break
            issue(nxt, host); pending.append(nxt); nxt = (nxt + 1) % self.depth
        while pending:
            slot = pending.pop(0)
            host = next(it, None)
            if host is not None:                             # keep the pipe full before handing out
                issue(nxt, host); pending.append(nxt); nxt = (nxt + 1) % self.depth
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ready[slot])
            yield ring[slot]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            freed[slot] = ev
