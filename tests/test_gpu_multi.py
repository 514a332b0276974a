"""Two-rank (2 x B200, NCCL) parity of APTS_D / APTS_P against goldens produced by the
unmodified reference on 2 gloo ranks.  Skipped on boxes with fewer than 2 GPUs."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, ws, port, case, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import torch.nn as nn
    from _helpers import FFNN, load_golden, set_flat
    import dd4ml_b200
    from dd4ml_b200 import APTS_D, APTS_P
    from test_gpu_optimizers import cfg

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws)
    name, kind, glob, loc, so, dog = case
    g = load_golden(name)
    dev = f"cuda:{rank}"
    model = FFNN(36, [12, 12], 10)
    set_flat(model, torch.from_numpy(g["w0"]))
    model = model.to(dev)
    X, Y = torch.from_numpy(g["X"]), torch.from_numpy(g["Y"])
    if kind == "d":
        per = X.shape[0] // ws
        Xr, Yr = X[rank * per:(rank + 1) * per].to(dev), Y[rank * per:(rank + 1) * per].to(dev)
        c = cfg(glob_second_order=so, glob_dogleg=dog, loc_second_order=so, loc_dogleg=dog)
        cls, extra = APTS_D, dict(foc=True)
    else:
        Xr, Yr = X.to(dev), Y.to(dev)
        c = cfg(delta=0.1, min_delta=0.01, max_delta=0.5, norm_type=math.inf, glob_second_order=so,
                loc_second_order=so)
        cls, extra = APTS_P, {}
    c.glob_opt, c.loc_opt = glob, loc
    c = cls.setup_APTS_hparams(c)
    if kind == "p":
        # reproduce the reference's ownership (the goldens record it) instead of a fresh randperm
        import dd4ml_b200.optimizers.apts_p as ap
        owned = set(int(i) for i in g[f"owned_tensors_rank{rank}"])

        def mark(m):
            for i, p in enumerate(m.parameters()):
                p.requires_grad = i in owned
            m._trainable_indices = sorted(owned)
            return m
        ap.mark_trainable = mark
    opt = cls(model.parameters(), model=model, criterion=nn.CrossEntropyLoss(), device=dev, nr_models=ws,
              glob_opt=c.glob_opt, glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt,
              loc_opt_hparams=c.loc_opt_hparams, glob_pass=True, norm_type=c.norm_type, max_loc_iters=3,
              max_glob_iters=1, tol=1e-6, **extra, **c.apts_params)
    losses, deltas, gev = [], [], []
    for _ in range(int(g["steps"])):
        losses.append(float(opt.step(Xr, Yr)))
        deltas.append(float(opt.delta))
        gev.append(float(opt.grad_evals))
    w = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    q.put((rank, dict(loss=losses, delta=deltas, gev=gev, w=w)))
    dist.barrier()
    dist.destroy_process_group()


CASES = [
    ("apts_d_tr_fo_ws2", "d", "tr", "tr", False, False),
    ("apts_d_tr_so_ws2", "d", "tr", "tr", True, False),
    ("apts_d_lssr1_so_ws2", "d", "lssr1_tr", "lssr1_tr", True, True),
    ("apts_p_tr_fo_ws2", "p", "tr", "tr", False, False),
    ("apts_p_tr_so_ws2", "p", "tr", "tr", True, False),
]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_two_rank_apts_matches_reference(case):
    from _helpers import load_golden
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + CASES.index(case)
    procs = [ctx.Process(target=_worker, args=(r, ws, port, case, q)) for r in range(ws)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=600) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
    g = load_golden(case[0])
    assert np.array_equal(outs[0]["w"], outs[1]["w"]), "global model differs between ranks"
    so = case[4]
    k = 4 if so else int(g["steps"])
    np.testing.assert_allclose(outs[0]["loss"][:k], g["loss"][:k], rtol=1e-3)
    if not so:
        np.testing.assert_allclose(outs[0]["delta"], g["delta"], rtol=1e-9)
        np.testing.assert_allclose(outs[0]["gev"], g["grad_evals"], atol=1e-9)
    assert outs[0]["loss"][-1] < outs[0]["loss"][0]
