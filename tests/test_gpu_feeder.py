"""GPU-resident batch feed: the row-gather kernel against torch indexing, and
an APTS_D run fed from it against the same run fed with host-collated batches."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from _helpers import FFNN, load_golden, set_flat

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from dd4ml_b200 import APTS_D, _lib
    from dd4ml_b200.dataloaders import DeviceBatchFeeder, MicroBatchFlattenSampler, MicroBatchOverlapSampler, OverlapBatchSampler

DEV = "cuda"


@pytest.mark.parametrize("shape,dtype", [((1, 28, 28), torch.float32), ((3,), torch.float32), ((5,), torch.float64),
                                         ((), torch.int64), ((7,), torch.uint8), ((3, 32, 32), torch.float32)])
def test_gather_rows_matches_torch_indexing(shape, dtype):
    gen = torch.Generator().manual_seed(0)
    n = 1000
    if dtype.is_floating_point:
        X = torch.randn((n,) + shape, generator=gen).to(dtype)
    else:
        X = torch.randint(0, 100, (n,) + shape, generator=gen).to(dtype)
    order = torch.randperm(n, generator=gen).tolist()
    sampler = OverlapBatchSampler(order, 96, overlap=0.25, drop_last=False)
    feeder = DeviceBatchFeeder([X], sampler, device=DEV)
    batches = list(sampler)
    assert len(feeder) == len(batches)
    for i, idx in enumerate(batches):
        (got,) = feeder.batch(i)
        assert got.dtype == X.dtype and tuple(got.shape) == (len(idx),) + shape
        assert torch.equal(got.cpu(), X[idx])
    # empty and ragged cases
    empty = DeviceBatchFeeder([X], [], device=DEV)
    assert len(empty) == 0 and list(empty) == []
    with pytest.raises(IndexError):
        DeviceBatchFeeder([X], [[0, n]], device=DEV)
    with pytest.raises(ValueError):
        DeviceBatchFeeder([X, X[:10]], [[0]], device=DEV)


def test_apts_d_fed_from_device_matches_host_batches():
    g = load_golden("apts_d_tr_fo_ws1")
    X, Y = torch.from_numpy(g["X"]), torch.from_numpy(g["Y"])
    n = X.shape[0]
    ob = OverlapBatchSampler(list(range(n)), 32, overlap=0.25, drop_last=True)
    sampler = MicroBatchFlattenSampler(MicroBatchOverlapSampler(ob, 2))

    def run(feed):
        from test_gpu_optimizers import _apts, cfg
        model = FFNN(36, [12, 12], 10)
        set_flat(model, torch.from_numpy(g["w0"]))
        opt = _apts("d", model.to(DEV), "tr", "tr", cfg(), foc=True)
        return [float(opt.step(xb, yb)) for xb, yb in feed]

    feeder = DeviceBatchFeeder([X, Y], sampler, device=DEV)
    l0 = _lib.get().launches
    dev_losses = run(feeder)
    assert _lib.get().launches - l0 >= 2 * len(feeder)
    host_losses = run((X[idx].to(DEV), Y[idx].to(DEV)) for idx in sampler)
    assert len(dev_losses) == len(host_losses) == len(feeder) >= 4
    assert dev_losses == host_losses


def test_apts_d_fed_through_prefetcher_ring_matches_resident_batches():
    """The closure graphs read the prefetcher's two ring buffers in place (one captured graph per buffer
    address, dd4ml_b200/graphs.py): different batch contents at a recurring address must be seen."""
    from dd4ml_b200.dataloaders import HostBatchPrefetcher
    from dd4ml_b200.graphs import GraphedEval
    g = load_golden("apts_d_tr_fo_ws1")
    X, Y = torch.from_numpy(g["X"]), torch.from_numpy(g["Y"])
    gen = torch.Generator().manual_seed(3)
    idx = [torch.randperm(X.shape[0], generator=gen)[:48] for _ in range(7)]
    host = [(X[i].clone().pin_memory(), Y[i].clone().pin_memory()) for i in idx]

    def run(feed):
        from test_gpu_optimizers import _apts, cfg
        model = FFNN(36, [12, 12], 10)
        set_flat(model, torch.from_numpy(g["w0"]))
        opt = _apts("d", model.to(DEV), "tr", "tr", cfg(), foc=True)
        losses = [float(opt.step(xb, yb)) for xb, yb in feed]
        evs = [v for v in vars(opt).values() if isinstance(v, GraphedEval)]
        return losses, evs

    ring_losses, evs = run(HostBatchPrefetcher(host, DEV, depth=2))
    res_losses, _ = run((x.to(DEV), y.to(DEV)) for x, y in host)
    assert ring_losses == res_losses and len(ring_losses) == 7
    for ev in evs:                                   # two addresses -> two in-place graphs, no private copy
        for ents in ev.cache.values():
            assert 1 <= len(ents) <= 2 and not any(e.private for e in ents)


def test_closure_graph_cache_is_bounded_over_changing_batch_sizes(monkeypatch):
    """The Trainer's batch-size growth loop (trainer.py:516-566) keeps producing new batch shapes: the closure
    graphs of old shapes are dropped (at most graphs.MAX_SHAPES kept) and results equal the eager closures."""
    from dd4ml_b200 import graphs
    g = load_golden("apts_d_tr_fo_ws1")
    X, Y = torch.from_numpy(g["X"]).to(DEV), torch.from_numpy(g["Y"]).to(DEV)
    sizes = [32, 40, 48, 56, 64, 72, 32, 72]

    def run():
        from test_gpu_optimizers import _apts, cfg
        model = FFNN(36, [12, 12], 10)
        set_flat(model, torch.from_numpy(g["w0"]))
        opt = _apts("d", model.to(DEV), "tr", "tr", cfg(), foc=True)
        return [float(opt.step(X[:b], Y[:b])) for b in sizes], opt

    graphed, opt = run()
    for ev in (v for v in vars(opt).values() if isinstance(v, graphs.GraphedEval)):
        assert 1 <= len(ev.cache) <= graphs.MAX_SHAPES and ev.captures >= len(set(sizes))
    monkeypatch.setattr(graphs, "ENABLED", False)
    eager, _ = run()
    assert graphed == eager


def test_host_batch_prefetcher_delivers_every_batch_in_order():
    from dd4ml_b200.dataloaders import HostBatchPrefetcher
    gen = torch.Generator().manual_seed(1)
    host = [(torch.randn(257, 33, generator=gen).pin_memory(), torch.randint(0, 9, (257,), generator=gen).pin_memory())
            for _ in range(7)]
    seen = []
    for xb, yb in HostBatchPrefetcher(host, DEV, depth=2):
        assert xb.is_cuda and yb.is_cuda
        # consume on the compute stream with some work in between, as an optimizer step would
        seen.append((xb.clone(), yb.clone()))
        torch.mm(torch.ones(512, 512, device=DEV), torch.ones(512, 512, device=DEV))
    assert len(seen) == len(host)
    for (xd, yd), (xh, yh) in zip(seen, host):
        assert torch.equal(xd.cpu(), xh) and torch.equal(yd.cpu(), yh)
    assert list(HostBatchPrefetcher([], DEV)) == []
    assert len(list(HostBatchPrefetcher(host[:1], DEV, depth=3))) == 1
