#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (cruzas/DD4ML).

Run in the authoring container only (needs /root/reference):

    OMP_NUM_THREADS=1 python tests/golden/gen_golden.py

The reference cannot travel to the GPU box, so the outputs (tests/golden/*.npz)
are committed.  Everything here drives the reference's own classes
(`dd4ml.optimizers.{tr,lssr1_tr,apts_d,apts_p,apts_pinn,tradam}`) through their public
API on synthetic tensors with fixed seeds; nothing is re-implemented.
Recorded: torch version, thread count, seeds, per-step scalars, and for the first
few sub-problem solves the full input/output vectors.
"""
import math
import os
import sys

os.environ.setdefault("OMP_NUM_THREADS", "1")
sys.path.insert(0, "/root/reference/src")

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import dd4ml.utility  # noqa: F401  (must precede dd4ml.optimizers: circular import)
from dd4ml.models.ffnn.medium_ffnn import MediumFFNN
from dd4ml.optimizers import lssr1_tr as ref_lssr1_mod
from dd4ml.optimizers import tr as ref_tr_mod
from dd4ml.optimizers.apts_d import APTS_D
from dd4ml.optimizers.apts_p import APTS_P
from dd4ml.optimizers.apts_pinn import APTS_PINN
from dd4ml.optimizers.lsr1 import LSR1
from dd4ml.optimizers.lssr1_tr import LSSR1_TR
from dd4ml.optimizers.tr import TR
from dd4ml.solvers.obs import OBS
from dd4ml.utility import CfgNode, set_seed
from dd4ml.utility.optimizer_utils import solve_tr_second_order as ref_solve_so

HERE = os.path.dirname(os.path.abspath(__file__))
IN_F, HID, OUT = 36, [12, 12], 10          # tiny FFNN: n = 730
MAX_SOLVES = 8


def make_model(dtype=torch.float32):
    set_seed(42)
    cfg = MediumFFNN.get_default_config()
    cfg.input_features, cfg.fc_layers, cfg.output_classes = IN_F, list(HID), OUT
    return MediumFFNN(cfg).to(dtype)


def make_data(n_samples, seed=1234, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n_samples, 1, 6, 6, generator=g).to(dtype)
    Y = torch.randint(0, OUT, (n_samples,), generator=g)
    return X, Y


def flat(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()


# ------------------------------------------------------------------ capture
class SolveRecorder:
    """Wraps solve_tr_second_order / OBS.Newton as imported into the optimizer modules."""

    def __init__(self):
        self.recs = []
        self._newton = []
        self._orig_newton = OBS.Newton

    def __enter__(self):
        rec = self

        def newton(obs, x0, Lam, a, delta):
            out = rec._orig_newton(obs, x0, Lam, a, delta)
            rec._newton.append((float(x0), float(out), Lam.clone(), a.clone()))
            return out

        def solve(gradient, grad_norm, trust_radius, lsr1_hessian, obs_solver, tol, dogleg=False):
            rec._newton.clear()
            p, pred = ref_solve_so(gradient=gradient, grad_norm=grad_norm, trust_radius=trust_radius,
                                   lsr1_hessian=lsr1_hessian, obs_solver=obs_solver, tol=tol, dogleg=dogleg)
            if len(rec.recs) < MAX_SOLVES:
                k = lsr1_hessian._current_pairs
                r = dict(g=gradient.clone(), gn=float(grad_norm), delta=float(trust_radius),
                         gamma=float(lsr1_hessian.gamma), S=lsr1_hessian._S_matrix[:, :k].clone(),
                         Y=lsr1_hessian._Y_matrix[:, :k].clone(), p=p.clone(), pred=float(pred),
                         dogleg=int(dogleg), tol=float(tol), newton=int(len(rec._newton) > 0))
                if rec._newton:
                    x0, sig, Lam, a = rec._newton[-1]
                    r.update(x0=x0, sigma=sig, Lam=Lam, a=a)
                rec.recs.append(r)
            return p, pred

        OBS.Newton = newton
        ref_tr_mod.solve_tr_second_order = solve
        ref_lssr1_mod.solve_tr_second_order = solve
        return self

    def __exit__(self, *exc):
        OBS.Newton = self._orig_newton
        ref_tr_mod.solve_tr_second_order = ref_solve_so
        ref_lssr1_mod.solve_tr_second_order = ref_solve_so

    def dump(self, out):
        out["n_solves"] = np.int64(len(self.recs))
        for i, r in enumerate(self.recs):
            for k, v in r.items():
                out[f"solve{i}_{k}"] = v.numpy() if torch.is_tensor(v) else np.asarray(v)


def meta(out, **kw):
    out["torch_version"] = np.asarray(torch.__version__)
    out["omp_threads"] = np.int64(torch.get_num_threads())
    for k, v in kw.items():
        out[k] = np.asarray(v)


def save(name, out):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KB)")


# ------------------------------------------------------------------ scenarios
def tr_kwargs(second_order, dogleg, norm_type=2):
    cfg = CfgNode(delta=0.1, max_delta=2.0, min_delta=1e-3, norm_type=norm_type,
                  glob_second_order=second_order, glob_dogleg=dogleg, tol=1e-6)
    from dd4ml.utility import get_tr_hparams
    return get_tr_hparams(cfg)


def run_plain(name, make_opt, steps, dtype=torch.float32, data_seed=1234):
    """Plain TR / LSSR1_TR as the top-level optimizer with an own closure."""
    model = make_model(dtype)
    X, Y = make_data(256, seed=data_seed, dtype=dtype)
    crit = nn.CrossEntropyLoss()
    opt = make_opt(model)

    def closure(compute_grad=True):
        opt.zero_grad()
        loss = crit(model(X), Y)
        if compute_grad:
            loss.backward()
        return loss

    out = {"w0": flat(model).numpy(), "X": X.numpy(), "Y": Y.numpy()}
    losses, deltas, ws = [], [], []
    with SolveRecorder() as rec:
        for _ in range(steps):
            loss, _g = opt.step(closure)
            losses.append(float(loss))
            deltas.append(float(opt.delta))
            ws.append(flat(model).numpy())
        rec.dump(out)
    out["loss"] = np.asarray(losses)
    out["delta"] = np.asarray(deltas)
    out["w_step1"] = ws[0]
    out["w_step5"] = ws[4]
    out["w_final"] = ws[-1]
    if getattr(opt, "hess", None) is not None:
        out["final_pairs"] = np.int64(len(opt.hess._S))
        out["final_gamma"] = np.float64(float(opt.hess.gamma))
    meta(out, steps=steps)
    save(name, out)


def apts_cfg(glob, loc, so, dog, m, norm_type=2, delta=0.1, min_delta=0.01, max_delta=1.0):
    return CfgNode(delta=delta, min_delta=min_delta, max_delta=max_delta, norm_type=norm_type,
                   glob_second_order=so, glob_dogleg=dog, loc_second_order=so, loc_dogleg=dog,
                   mem_length=m, max_wolfe_iters=5, max_zoom_iters=5, tol=1e-6,
                   paper_tr_update=False, glob_opt=glob, loc_opt=loc, learning_rate=0.1)


def _apts_worker(rank, ws, port, kind, glob, loc, so, dog, m, steps, dtype, q):
    torch.set_num_threads(1)
    if ws > 1 or kind == "apts_p":
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=ws)
    model = make_model(dtype)
    per = 128
    X, Y = make_data(per * ws, dtype=dtype)
    crit = nn.CrossEntropyLoss()
    norm_type = math.inf if kind == "apts_p" else 2
    kw = dict(delta=0.1, min_delta=0.01, max_delta=0.5) if kind == "apts_p" else {}
    cfg = apts_cfg(glob, loc, so, dog, m, norm_type=norm_type, **kw)
    cls = APTS_D if kind == "apts_d" else APTS_P
    cfg = cls.setup_APTS_hparams(cfg)
    extra = dict(foc=True) if kind == "apts_d" else {}
    opt = cls(model.parameters(), model=model, criterion=crit, device="cpu", nr_models=ws,
              glob_opt=cfg.glob_opt, glob_opt_hparams=cfg.glob_opt_hparams, loc_opt=cfg.loc_opt,
              loc_opt_hparams=cfg.loc_opt_hparams, glob_pass=True, norm_type=norm_type,
              max_loc_iters=3, max_glob_iters=1, tol=1e-6, **extra, **cfg.apts_params)
    if kind == "apts_d":
        Xr, Yr = X[rank * per:(rank + 1) * per], Y[rank * per:(rank + 1) * per]
    else:
        Xr, Yr = X, Y           # APTS_P: every rank evaluates the same batch
    out = {"w0": flat(model).numpy(), "X": X.numpy(), "Y": Y.numpy()}
    if kind == "apts_p":
        out["owned_tensors"] = np.asarray(sorted(int(i) for i in opt.loc_model._trainable_indices))
    losses, deltas, gdeltas, ldeltas, gevals, ws_hist = [], [], [], [], [], []
    for _ in range(steps):
        loss = opt.step(Xr, Yr)
        losses.append(float(loss))
        deltas.append(float(opt.delta))
        gdeltas.append(float(opt.glob_opt.delta))
        ldeltas.append(float(opt.loc_opt.delta))
        gevals.append(float(opt.grad_evals))
        ws_hist.append(flat(model).numpy())
    out.update(loss=np.asarray(losses), delta=np.asarray(deltas), glob_delta=np.asarray(gdeltas),
               loc_delta=np.asarray(ldeltas), grad_evals=np.asarray(gevals),
               w_step1=ws_hist[0], w_step3=ws_hist[2], w_final=ws_hist[-1])
    q.put((rank, out))
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def run_apts(name, kind, ws, glob, loc, so, dog, m, steps, dtype=torch.float32, port=29611):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_apts_worker,
                         args=(r, ws, port, kind, glob, loc, so, dog, m, steps, dtype, q))
             for r in range(ws)]
    for p in procs:
        p.start()
    outs = dict(q.get() for _ in range(ws))
    for p in procs:
        p.join()
    out = outs[0]
    if kind == "apts_p":
        for r in range(ws):
            out[f"owned_tensors_rank{r}"] = outs[r]["owned_tensors"]
    for r in range(1, ws):      # the global model must be identical on every rank
        assert np.array_equal(outs[r]["w_final"], out["w_final"]), "ranks diverged"
    meta(out, steps=steps, world_size=ws, glob=glob, loc=loc, second_order=int(so),
         dogleg=int(dog), mem_length=m, kind=kind)
    save(name, out)


class SineCrit(nn.Module):
    """PINN-like criterion for APTS_PINN: reads the collocation points from
    `current_x` (set by APTS_PINN._set_criterion_input), ignores the (long) labels."""
    low, high = 0.0, 1.0

    def __init__(self):
        super().__init__()
        self.current_x = None

    def forward(self, out, labels):
        return ((out.squeeze(-1) - torch.sin(2 * math.pi * self.current_x.squeeze(-1))) ** 2).mean()


def _pinn_worker(rank, ws, port, steps, q):
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    set_seed(42)
    cfg = MediumFFNN.get_default_config()
    cfg.input_features, cfg.fc_layers, cfg.output_classes = 1, [16, 16], 1

    class Net(MediumFFNN):          # same stack, linear head instead of log_softmax
        def forward(self, x):
            x = x.view(x.size(0), -1)
            for i in range(1, len(self.config.fc_layers) + 1):
                x = getattr(self, f"stage{i}")(x)
            return self.finish(x)

    globals()["Net"] = Net
    model = Net(cfg)
    crit = SineCrit()
    x = torch.linspace(0, 1, 65).unsqueeze(1)
    lab = torch.zeros(65, 1)
    tr_kw = dict(delta=0.1, nu_dec=0.25, nu_inc=0.75, inc_factor=1.2, dec_factor=0.9,
                 max_delta=2.0, min_delta=1e-3, tol=1e-6)
    opt = APTS_PINN(model.parameters(), model=model, criterion=crit, device="cpu", glob_opt=TR,
                    glob_opt_hparams=tr_kw, loc_opt=TR, loc_opt_hparams=tr_kw, glob_pass=True,
                    foc=True, norm_type=2, max_loc_iters=4, max_glob_iters=1, num_subdomains=ws,
                    overlap=0.1, **tr_kw)
    out = {"w0": flat(model).numpy(), "x": x.numpy()}
    losses, deltas = [], []
    for _ in range(steps):
        loss = opt.step(inputs=x, labels=lab)
        losses.append(float(loss))
        deltas.append(float(opt.delta))
    out.update(loss=np.asarray(losses), delta=np.asarray(deltas), w_final=flat(model).numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def run_pinn(name, ws, steps, port=29631):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pinn_worker, args=(r, ws, port, steps, q)) for r in range(ws)]
    for p in procs:
        p.start()
    outs = dict(q.get() for _ in range(ws))
    for p in procs:
        p.join()
    out = outs[0]
    meta(out, steps=steps, world_size=ws, overlap=0.1)
    save(name, out)


# ------------------------------------------------------------------ BASELINE config 5 as the Trainer runs it
def _config5_worker(rank, ws, port, glob, loc, paper, delta, min_delta, max_delta, steps, q):
    """PINNFFNN + AllenCahnPINNLoss + APTS_P exactly as `Trainer.run_by_epoch_PINN` /
    `_train_one_batch_PINN` drive them (trainer.py:985-1002, 1203-1214): everything is BUILT in
    float32, then `model` and `optimizer.loc_model` are converted to float64 and the optimizer is
    stepped with double inputs that require grad and the boundary mask as (double) labels."""
    from dd4ml.datasets.pinn_allencahn import AllenCahn1DDataset
    from dd4ml.models.ffnn.pinn_ffnn import PINNFFNN
    from dd4ml.utility.pinn_allencahn_loss import AllenCahnPINNLoss
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    set_seed(42)
    ds = AllenCahn1DDataset(AllenCahn1DDataset.get_default_config())
    model = PINNFFNN(PINNFFNN.get_default_config())
    crit = AllenCahnPINNLoss()
    cfg = CfgNode(delta=delta, min_delta=min_delta, max_delta=max_delta, norm_type=math.inf,
                  glob_second_order=False, glob_dogleg=False, loc_second_order=False, loc_dogleg=False,
                  mem_length=5, max_wolfe_iters=5, max_zoom_iters=5, tol=1e-6, paper_tr_update=paper,
                  glob_opt=glob, loc_opt=loc, learning_rate=0.1)
    cfg = APTS_P.setup_APTS_hparams(cfg)
    opt = APTS_P(model.parameters(), model=model, criterion=crit, device="cpu", nr_models=ws,
                 glob_opt=cfg.glob_opt, glob_opt_hparams=cfg.glob_opt_hparams, loc_opt=cfg.loc_opt,
                 loc_opt_hparams=cfg.loc_opt_hparams, glob_pass=True, norm_type=math.inf, max_loc_iters=3,
                 max_glob_iters=1, tol=1e-6, **cfg.apts_params)
    out = {"w0_f32": flat(model).numpy()}
    model = model.double()                                   # trainer.py:1203
    opt.loc_model = opt.loc_model.double()                   # trainer.py:1211
    x = ds.data.double()                                     # trainer.py:987-989
    y = ds.boundary_mask.double()
    x.requires_grad_(True)
    crit.current_x = x                                       # trainer.py:997-998
    out.update(x=x.detach().numpy(), y=y.numpy(),
               owned_tensors=np.asarray(sorted(int(i) for i in opt.loc_model._trainable_indices)))
    losses, deltas, gdeltas, ldeltas, gevals, ws_hist = [], [], [], [], [], []
    for _ in range(steps):
        loss = opt.step(inputs=x, labels=y)
        losses.append(float(loss))
        deltas.append(float(opt.delta))
        gdeltas.append(float(opt.glob_opt.delta))
        ldeltas.append(float(opt.loc_opt.delta))
        gevals.append(float(opt.grad_evals))
        ws_hist.append(flat(model).numpy())
    out.update(loss=np.asarray(losses), delta=np.asarray(deltas), glob_delta=np.asarray(gdeltas),
               loc_delta=np.asarray(ldeltas), grad_evals=np.asarray(gevals),
               w_step1=ws_hist[0], w_step3=ws_hist[2], w_final=ws_hist[-1])
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def run_config5(name, ws, glob, loc, paper, steps, port, delta=0.1, min_delta=0.01, max_delta=0.5):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_config5_worker,
                         args=(r, ws, port, glob, loc, paper, delta, min_delta, max_delta, steps, q))
             for r in range(ws)]
    for p in procs:
        p.start()
    outs = dict(q.get() for _ in range(ws))
    for p in procs:
        p.join()
    out = outs[0]
    for r in range(ws):
        out[f"owned_tensors_rank{r}"] = outs[r]["owned_tensors"]
    for r in range(1, ws):
        assert np.array_equal(outs[r]["w_final"], out["w_final"]), "ranks diverged"
    meta(out, steps=steps, world_size=ws, glob=glob, loc=loc, paper_tr_update=int(paper), delta0=delta,
         min_delta=min_delta, max_delta=max_delta, kind="apts_p_config5")
    save(name, out)


def _pinn_literal_worker(rank, ws, port, steps, q):
    """tests/test_pinn_subdomains.py:141-204 (`_run_apts_pinn`) literally, plus a seed."""
    from dd4ml.datasets.pinn_allencahn import AllenCahn1DDataset
    from dd4ml.models.ffnn.pinn_ffnn import PINNFFNN
    from dd4ml.utility.pinn_allencahn_loss import AllenCahnPINNLoss
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    set_seed(42)
    ds = AllenCahn1DDataset(AllenCahn1DDataset.get_default_config())
    model = PINNFFNN(PINNFFNN.get_default_config())
    crit = AllenCahnPINNLoss()
    tr_kw = dict(delta=0.1, nu_dec=0.25, nu_inc=0.75, inc_factor=1.2, dec_factor=0.9, max_delta=2.0,
                 min_delta=1e-3, tol=1e-6)
    opt = APTS_PINN(model.parameters(), model=model, criterion=crit, device="cpu", glob_opt=TR,
                    glob_opt_hparams=tr_kw, loc_opt=TR, loc_opt_hparams=tr_kw, glob_pass=False, foc=False,
                    norm_type=2, max_loc_iters=10, max_glob_iters=1, num_subdomains=ws, **tr_kw)
    x = ds.data
    flag = torch.cat([torch.zeros(len(ds.x_interior), 1), torch.ones(len(ds.x_boundary), 1)])
    crit.current_x = x
    out = {"w0": flat(model).numpy(), "x": x.numpy(), "flag": flag.numpy()}
    losses, deltas, gevals, ws_hist = [], [], [], []
    for _ in range(steps):
        loss = opt.step(inputs=x, labels=flag)
        losses.append(float(loss))
        deltas.append(float(opt.delta))
        gevals.append(float(opt.grad_evals))
        ws_hist.append(flat(model).numpy())
    out.update(loss=np.asarray(losses), delta=np.asarray(deltas), grad_evals=np.asarray(gevals),
               w_step1=ws_hist[0], w_step3=ws_hist[2], w_final=ws_hist[-1])
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def run_pinn_literal(name, ws, steps, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pinn_literal_worker, args=(r, ws, port, steps, q)) for r in range(ws)]
    for p in procs:
        p.start()
    outs = dict(q.get() for _ in range(ws))
    for p in procs:
        p.join()
    out = outs[0]
    for r in range(1, ws):
        assert np.array_equal(outs[r]["w_final"], out["w_final"]), "ranks diverged"
    meta(out, steps=steps, world_size=ws, kind="apts_pinn_literal")
    save(name, out)


def run_config5_all():
    run_config5("config5_apts_p_tr_ws2", 2, "tr", "tr", False, 12, 29641)
    run_config5("config5_apts_p_lssr1_ws2", 2, "lssr1_tr", "lssr1_tr", True, 12, 29642)
    run_config5("config5_apts_p_tr_small_ws2", 2, "tr", "tr", False, 12, 29643, delta=0.002, min_delta=1e-5,
                max_delta=0.5)
    run_pinn_literal("apts_pinn_literal_ws2", 2, 10, 29644)


def run_kernel_level(name, ms=(1, 3, 5), seed=7):
    """Direct calls of LSR1.update_memory / precompute / B and solve_tr_second_order on
    synthetic (s, y) pairs: y = A s for a fixed indefinite diagonal-plus-low-rank A."""
    torch.manual_seed(seed)
    n = 512
    diag = torch.linspace(-0.5, 3.0, n)
    Ulr = torch.randn(n, 4) / math.sqrt(n)
    A = lambda v: diag * v + Ulr @ (Ulr.t() @ v)
    out = {}
    idx = 0
    for m in ms:
        hess = LSR1(gamma=1.0, memory_length=m, tol=1e-6)
        accepted = []
        for j in range(m + 2):
            s = torch.randn(n) * 0.05
            y = A(s) + 1e-3 * torch.randn(n)
            before = len(hess._S)
            g_before = float(hess.gamma)
            hess.update_memory(s, y)
            accepted.append(int(float(hess.gamma) != g_before or len(hess._S) != before))
            hess.precompute()
            out[f"upd{idx}_s"], out[f"upd{idx}_y"] = s.numpy(), y.numpy()
            out[f"upd{idx}_m"], out[f"upd{idx}_accepted"] = np.int64(m), np.int64(accepted[-1])
            out[f"upd{idx}_gamma"] = np.float64(float(hess.gamma))
            out[f"upd{idx}_Minv"] = hess.Minv.numpy().copy()
            v = torch.randn(n)
            out[f"upd{idx}_v"], out[f"upd{idx}_Bv"] = v.numpy(), hess.B(v).numpy()
            idx += 1
        for dog in (False, True):
            for delta in (0.01, 0.3, 5.0, 100.0):
                g = torch.randn(n)
                obs = OBS()
                newt = []
                orig = OBS.Newton

                def newton(o, x0, Lam, a, d, _n=newt):
                    r = orig(o, x0, Lam, a, d)
                    _n.append((float(x0), float(r), Lam.clone(), a.clone()))
                    return r
                OBS.Newton = newton
                p, pred = ref_solve_so(g, torch.norm(g), delta, hess, obs, 1e-6, dogleg=dog)
                OBS.Newton = orig
                key = f"so_m{m}_dog{int(dog)}_d{delta}"
                k = hess._current_pairs
                out[key + "_g"], out[key + "_p"] = g.numpy(), p.numpy()
                out[key + "_pred"] = np.float64(float(pred))
                out[key + "_S"] = hess._S_matrix[:, :k].numpy().copy()
                out[key + "_Y"] = hess._Y_matrix[:, :k].numpy().copy()
                out[key + "_gamma"] = np.float64(float(hess.gamma))
                out[key + "_newton"] = np.int64(len(newt))
                if newt:
                    out[key + "_x0"], out[key + "_sigma"] = np.float64(newt[-1][0]), np.float64(newt[-1][1])
                    out[key + "_Lam"], out[key + "_a"] = newt[-1][2].numpy(), newt[-1][3].numpy()
    out["n_updates"] = np.int64(idx)
    meta(out)
    save(name, out)


def run_tradam(name, norm_type, lr, steps, dtype=torch.float32):
    """TRAdam (optimizers/tradam.py) as a standalone optimizer; the closure back-propagates."""
    from dd4ml.optimizers.tradam import TRAdam
    model = make_model(dtype)
    X, Y = make_data(256, dtype=dtype)
    crit = nn.CrossEntropyLoss()
    opt = TRAdam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, norm_type=norm_type)

    def closure():
        opt.zero_grad()
        loss = crit(model(X), Y)
        loss.backward()
        return loss

    out = {"w0": flat(model).numpy(), "X": X.numpy(), "Y": Y.numpy()}
    losses, ws, smax = [], [], []
    for _ in range(steps):
        losses.append(float(opt.step(closure)))
        ws.append(flat(model).numpy())
        smax.append(float(opt._step_buf.abs().max()))
    out["loss"] = np.asarray(losses)
    out["step_max"] = np.asarray(smax)
    out["w_step1"], out["w_step5"], out["w_final"] = ws[0], ws[4], ws[-1]
    out["m_final"], out["v_final"] = opt._m_buf.numpy().copy(), opt._v_buf.numpy().copy()
    out["step_final"] = opt._step_buf.numpy().copy()
    meta(out, steps=steps, lr=lr)
    save(name, out)


def run_samplers(name):
    """Batch sequences of the reference's overlap / micro-batch samplers (dataloaders.py:18-229)."""
    import json
    from dd4ml.dataloaders import MicroBatchFlattenSampler, MicroBatchOverlapSampler, OverlapBatchSampler
    cases = []
    rng = np.random.default_rng(7)
    for n, bs, ov, drop, nsub in [(10, 4, 0.5, False, 2), (10, 4, 0.5, True, 2), (12, 4, 2, True, 3),
                                  (12, 4, 2, False, 3), (7, 3, 0.34, False, 2), (7, 3, 1, True, 2),
                                  (50, 16, 0.25, True, 4), (50, 16, 0.25, False, 4), (9, 5, 0.0, False, 0),
                                  (9, 5, 0, True, 0), (5, 8, 3, False, 2), (64, 8, 7, True, 8),
                                  (33, 10, 0.1, False, 5), (100, 25, 0.2, True, 2), (6, 2, 1, False, 3)]:
        order = rng.permutation(n).tolist()
        ob = OverlapBatchSampler(order, bs, overlap=ov, drop_last=drop)
        rec = {"n": n, "batch_size": bs, "overlap": ov, "drop_last": drop, "num_subdomains": nsub, "order": order,
               "overlap_sz": ob.overlap_sz, "len": len(ob), "batches": [list(map(int, b)) for b in ob]}
        if nsub > 0 and ob.overlap_sz > 0:
            try:
                mb = MicroBatchOverlapSampler(ob, nsub, allow_empty_microbatches=False)
                rec["micro"] = [[list(map(int, p)) for p in parts] for parts in mb]
                rec["flat"] = [list(map(int, p)) for p in MicroBatchFlattenSampler(mb)]
                rec["flat_len"] = len(MicroBatchFlattenSampler(mb))
                rec["info"] = mb.get_overlap_info()
            except (RuntimeError, ValueError) as e:
                rec["micro_error"] = type(e).__name__
        cases.append(rec)
    path = os.path.join(HERE, name + ".json")
    json.dump({"torch": torch.__version__, "cases": cases}, open(path, "w"))
    print(f"wrote {path}  ({os.path.getsize(path) // 1024} KB)")


def run_lssr1_m10():
    """LSSR1_TR with the reference's default memory length of 10 (second order + dogleg), 30 steps."""
    from dd4ml.utility import get_lssr1_tr_hparams
    hp = get_lssr1_tr_hparams(apts_cfg("lssr1_tr", "lssr1_tr", True, True, 10))
    hp["sync"] = False
    run_plain("lssr1_so_dog1_m10", lambda m: LSSR1_TR(m.parameters(), **hp), 30, data_seed=2010)


def main():
    if "--only-config5" in sys.argv:
        torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
        run_config5_all()
        return
    if "--only-tradam-loc" in sys.argv:
        torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
        run_apts("apts_d_tr_tradam_ws1", "apts_d", 1, "tr", "tradam", False, False, 3, 6, port=29661)
        return
    if "--only-dog1-seeds" in sys.argv:
        torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
        from dd4ml.utility import get_lssr1_tr_hparams
        hp = get_lssr1_tr_hparams(apts_cfg("lssr1_tr", "lssr1_tr", True, True, 3))
        hp["sync"] = False
        for j, seed in enumerate((2001, 2002, 2003), start=1):
            run_plain(f"lssr1_so_dog1_s{j}", lambda m, hp=hp: LSSR1_TR(m.parameters(), **hp), 30, data_seed=seed)
        return
    if "--only-ws8" in sys.argv:
        torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
        run_apts("apts_d_tr_fo_ws8", "apts_d", 8, "tr", "tr", False, False, 3, 8, port=29651)
        run_apts("apts_d_lssr1_fo_ws8", "apts_d", 8, "lssr1_tr", "lssr1_tr", False, False, 3, 8, port=29652)
        return
    if "--only-kernel-large" in sys.argv:
        torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
        run_kernel_level("kernel_level_m8_m10", ms=(8, 10), seed=11)
        run_lssr1_m10()
        return
    if "--only-samplers" in sys.argv:
        run_samplers("samplers")
        return
    if "--only-tradam" in sys.argv:
        torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
        run_tradam("tradam_inf", math.inf, 0.01, 20)
        run_tradam("tradam_l2", 2, 0.5, 20)
        return
    torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
    run_kernel_level("kernel_level")
    run_kernel_level("kernel_level_m8_m10", ms=(8, 10), seed=11)     # up to the reference's default memory length
    run_lssr1_m10()
    run_plain("tr_fo", lambda m: TR(m.parameters(), **tr_kwargs(False, False)), 30)
    run_plain("tr_fo_inf", lambda m: TR(m.parameters(), **tr_kwargs(False, False, math.inf)), 12)
    run_plain("tr_so", lambda m: TR(m.parameters(), **tr_kwargs(True, False)), 30)
    run_plain("tr_so_dogleg", lambda m: TR(m.parameters(), **tr_kwargs(True, True)), 30)
    from dd4ml.utility import get_lssr1_tr_hparams
    for dog in (False, True):
        hp = get_lssr1_tr_hparams(apts_cfg("lssr1_tr", "lssr1_tr", True, dog, 3))
        hp["sync"] = False
        run_plain(f"lssr1_so_dog{int(dog)}", lambda m, hp=hp: LSSR1_TR(m.parameters(), **hp), 30)
        if dog:
            for j, seed in enumerate((2001, 2002, 2003), start=1):
                run_plain(f"lssr1_so_dog1_s{j}", lambda m, hp=hp: LSSR1_TR(m.parameters(), **hp), 30, data_seed=seed)
    hp = get_lssr1_tr_hparams(apts_cfg("lssr1_tr", "lssr1_tr", False, False, 3))
    hp["sync"] = False
    run_plain("lssr1_fo", lambda m: LSSR1_TR(m.parameters(), **hp), 20)
    hp = dict(hp, paper_tr_update=True, second_order=True)
    run_plain("lssr1_paper", lambda m: LSSR1_TR(m.parameters(), **hp), 20)
    run_apts("apts_d_tr_fo_ws1", "apts_d", 1, "tr", "tr", False, False, 3, 12, port=29601)
    run_apts("apts_d_tr_fo_ws2", "apts_d", 2, "tr", "tr", False, False, 3, 12, port=29602)
    run_apts("apts_d_tr_so_ws2", "apts_d", 2, "tr", "tr", True, False, 3, 12, port=29603)
    run_apts("apts_d_lssr1_so_ws2", "apts_d", 2, "lssr1_tr", "lssr1_tr", True, True, 3, 12, port=29604)
    run_apts("apts_d_lssr1_fo_ws4", "apts_d", 4, "lssr1_tr", "lssr1_tr", False, False, 3, 8, port=29605)
    run_apts("apts_p_tr_fo_ws2", "apts_p", 2, "tr", "tr", False, False, 3, 12, port=29606)
    run_apts("apts_p_tr_fo_ws2_f64", "apts_p", 2, "tr", "tr", False, False, 3, 12,
             dtype=torch.float64, port=29607)
    run_apts("apts_p_tr_so_ws2", "apts_p", 2, "tr", "tr", True, False, 3, 12, port=29608)
    run_pinn("apts_pinn_ws2", 2, 10, port=29609)
    run_apts("apts_d_tr_fo_ws8", "apts_d", 8, "tr", "tr", False, False, 3, 8, port=29651)
    run_apts("apts_d_lssr1_fo_ws8", "apts_d", 8, "lssr1_tr", "lssr1_tr", False, False, 3, 8, port=29652)
    run_apts("apts_d_tr_tradam_ws1", "apts_d", 1, "tr", "tradam", False, False, 3, 6, port=29661)
    run_config5_all()
    run_tradam("tradam_inf", math.inf, 0.01, 20)
    run_tradam("tradam_l2", 2, 0.5, 20)
    run_samplers("samplers")


if __name__ == "__main__":
    main()
