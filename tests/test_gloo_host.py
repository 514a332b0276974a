"""world_size-2/3 `gloo` tests on CPU of the multi-rank side of the path (no GPU, no compute
kernels): the parameter-ownership logic of APTS_P (product host code, utility/apts_utils.py) and
the exchange step of APTS_D with REAL collectives -- one process per subdomain, the payloads laid
out as the product sends them over NCCL ([g || loss] averaged for every global closure,
[step || pred || evals] summed once per outer iteration) -- against the single-process oracle and
the golden of the unmodified reference (which itself ran on 2 gloo ranks).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    outs = {}
    for _ in range(world):
        r, out = q.get(timeout=300)
        outs[r] = out
    for p in procs:
        p.join(60)
    for r, o in outs.items():
        if isinstance(o, str) and o.startswith("ERROR"):
            raise AssertionError(f"rank {r}: {o}")
    return outs


def _entry(fn, rank, world, port, q, *args):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "ERROR " + traceback.format_exc()))


# ------------------------------------------------------------------ APTS_P ownership (product host logic)
def _ownership(rank, world, seed):
    from _helpers import FFNN
    from dd4ml_b200.utility.apts_utils import (mark_trainable, trainable_grads_to_vector,
                                               trainable_params_to_vector)
    torch.manual_seed(1000 + rank)                # different RNG streams: only rank 0's permutation may count
    model = FFNN(36, [12, 12, 12], 10)            # 8 parameter tensors
    torch.manual_seed(seed if rank == 0 else 7 * rank + 1)
    mark_trainable(model)
    owned = [int(i) for i in model._trainable_indices]
    flags = [bool(p.requires_grad) for p in model.parameters()]
    for p in model.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    nv = int(trainable_params_to_vector(model).numel())
    gv = trainable_grads_to_vector(model)
    return dict(owned=owned, flags=flags, nv=nv, gsum=float(gv.sum()), gnum=int(gv.numel()),
                numels=[p.numel() for p in model.parameters()])


@pytest.mark.parametrize("world", [2, 3])
def test_apts_p_parameter_ownership_is_a_partition(world):
    seed = 4242
    outs = _spawn(_ownership, world, seed)
    T = len(outs[0]["flags"])
    # the permutation every rank used is the one rank 0 drew (apts_utils.py:83-90)
    torch.manual_seed(seed)
    perm = torch.randperm(T, dtype=torch.long).tolist()
    q, r = divmod(T, world)
    start = 0
    seen = []
    for rank in range(world):
        size = q + (1 if rank < r else 0)
        assert outs[rank]["owned"] == perm[start:start + size]                   # apts_utils.py:93-98
        assert [i for i, f in enumerate(outs[rank]["flags"]) if f] == sorted(outs[rank]["owned"])
        numels = outs[rank]["numels"]
        assert outs[rank]["nv"] == outs[rank]["gnum"] == sum(numels[i] for i in outs[rank]["owned"])
        assert outs[rank]["gsum"] == pytest.approx((rank + 1) * outs[rank]["nv"])
        seen += outs[rank]["owned"]
        start += size
    assert sorted(seen) == list(range(T))                                        # disjoint cover


def _too_many_ranks(rank, world):
    import torch.nn as nn
    from dd4ml_b200.utility.apts_utils import mark_trainable
    try:
        mark_trainable(nn.Linear(3, 2, bias=False))      # one tensor, two ranks
    except ValueError as e:
        return str(e)
    return "no error"


def test_apts_p_more_ranks_than_tensors_raises_like_reference():
    outs = _spawn(_too_many_ranks, 2)
    assert all("World size cannot be greater" in o for o in outs.values())


# ------------------------------------------------------------------ APTS_D exchange with real collectives
def _apts_d_rank(rank, world, golden, steps, so):
    import oracle
    from _helpers import FFNN, load_golden, make_fg
    g = load_golden(golden)
    per = g["X"].shape[0] // world
    X = torch.from_numpy(g["X"])[rank * per:(rank + 1) * per]
    Y = torch.from_numpy(g["Y"])[rank * per:(rank + 1) * per]
    fg = make_fg(FFNN(36, [12, 12], 10), X, Y)
    w0 = torch.from_numpy(g["w0"])
    kw = dict(second_order=so, dogleg=False)
    o = oracle.APTSDOracle(w0, world, "tr", oracle.tr_hparams(0.1, 0.01, 1.0, **kw), "tr",
                           oracle.loc_tr_hparams(0.1, 0.01, 1.0, world, **kw), **oracle.apts_hparams(0.1, 0.01, 1.0),
                           glob_pass=True, foc=True, norm_type=2, max_loc_iters=3, max_glob_iters=1, tol=1e-6)
    n = w0.numel()
    loc_opt, w_loc = o.loc_opts[0], o.w_loc[0]          # this process owns ONE subdomain
    state = {"evals": 0.0, "loc_evals": 0.0, "resid": None}

    def fg_glob(w):                                      # [g || loss] in one all-reduce, averaged
        loss, grad = fg(w)
        buf = torch.cat([grad, loss.reshape(1)])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        buf /= world
        state["evals"] += 1.0
        return buf[n].clone(), buf[:n].clone()

    def fg_loc(w):                                       # apts_base.py:245-292 (FOC: value only)
        loss, grad = fg(w)
        state["loc_evals"] += 1.0
        diff = w - o.w
        if diff.norm() > o.tol and state["resid"] is not None:
            loss = loss + (state["resid"] @ diff)
        return loss, grad

    losses, deltas, gevals = [], [], []
    for _ in range(steps):
        state["evals"] = state["loc_evals"] = 0.0
        w_start = o.w.clone()
        f0, g0 = fg_glob(o.w)
        lf0, lg0 = fg_loc(w_loc)
        state["resid"] = g0 - lg0
        ll, _ = o._step_loop(loc_opt, fg_loc, w_loc, o.max_loc_iters, lf0, lg0)
        # additive Schwarz: [step || pred || evals] summed in ONE all-reduce (apts_d.py:114-151)
        buf = torch.cat([w_loc - w_start, (lf0 - ll).reshape(1), torch.tensor([state["loc_evals"]])])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        step, pred, evals = buf[:n].clone(), buf[n].clone(), float(buf[n + 1]) / world
        loss, grad, o.glob_opt.delta = o._control(fg_glob, w_start, f0, g0, step, pred)
        loss, grad = o._step_loop(o.glob_opt, fg_glob, o.w, o.max_glob_iters, loss, grad)
        o.delta = o.glob_opt.delta
        loc_opt.delta = o.glob_opt.delta / world
        w_loc.copy_(o.w)
        losses.append(float(loss)); deltas.append(float(o.delta)); gevals.append(state["evals"] + evals)
    return dict(loss=losses, delta=deltas, grad_evals=gevals, w=o.w.numpy().copy())


@pytest.mark.parametrize("golden,so", [("apts_d_tr_fo_ws2", False), ("apts_d_tr_so_ws2", True)])
def test_apts_d_exchange_over_gloo_matches_reference_golden(golden, so):
    from _helpers import load_golden
    g = load_golden(golden)
    steps = int(g["steps"])
    outs = _spawn(_apts_d_rank, 2, golden, steps, so)
    assert np.array_equal(outs[0]["w"], outs[1]["w"])                  # the global model is replicated
    assert outs[0]["loss"] == outs[1]["loss"] and outs[0]["delta"] == outs[1]["delta"]
    k = steps if not so else 4          # second order: until LAPACK's eigenvector signs matter
    np.testing.assert_allclose(outs[0]["loss"][:k], g["loss"][:k], rtol=2e-5)
    np.testing.assert_allclose(outs[0]["delta"][:k], g["delta"][:k], rtol=1e-9)
    np.testing.assert_allclose(outs[0]["grad_evals"][:k], g["grad_evals"][:k], atol=1e-9)
    if not so:
        assert float(np.linalg.norm(outs[0]["w"] - g["w_final"]) / np.linalg.norm(g["w_final"])) < 1e-5
