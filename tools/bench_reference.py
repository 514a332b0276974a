"""`bench.py --impl reference`: the UNMODIFIED reference (cruzas/DD4ML, installed in `baseline/_ref` by
`pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`)
driven through its own public API on the host cores of the box:

    dd4ml.optimizers.{apts_d.APTS_D, apts_p.APTS_P, tr.TR} + the reference's own model classes,
    one gloo process per subdomain (what tests/run_config_file.py spawns), OMP threads = cores / N.

None of dd4ml_b200's models, kernels or host code is on this path.  When `baseline/_ref` is missing
the caller falls back to the oracle port (kind "port").
"""
from __future__ import annotations

import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "dd4ml"))


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dd4ml.utility  # noqa: F401  (must precede dd4ml.optimizers: circular import)


def _apts_cfg(CfgNode, **kw):
    base = dict(delta=0.1, min_delta=0.01, max_delta=1.0, norm_type=2, glob_second_order=True, glob_dogleg=True,
                loc_second_order=True, loc_dogleg=True, mem_length=3, max_wolfe_iters=5, max_zoom_iters=5, tol=1e-6,
                paper_tr_update=False, glob_opt="lssr1_tr", loc_opt="lssr1_tr", learning_rate=0.1)
    base.update(kw)
    return CfgNode(**base)


def build(config: str, rank: int, ws: int, per_rank_batch: int, overrides: dict):
    """Returns step() -> loss (one unit of the metric) built from reference classes only."""
    import torch
    import torch.nn as nn
    _import_reference()
    from dd4ml.utility import CfgNode, set_seed
    sys.path.insert(0, ROOT)
    from dd4ml_b200 import workloads as W          # synthetic DATA generators only (no model / optimizer code)

    set_seed(42)
    if config in ("ffnn_apts_d", "cnn_apts_d", "gpt_apts_d"):
        from dd4ml.optimizers.apts_d import APTS_D
        if config == "ffnn_apts_d":
            from dd4ml.models.ffnn.medium_ffnn import MediumFFNN
            model = MediumFFNN(MediumFFNN.get_default_config())          # 784-[128]x8-10
            X, Y = W.mnist_shape_batch(per_rank_batch, 1234 + rank)
            if overrides.get("labels") == "teacher":
                Y = W.teacher_labels(X)
            crit = nn.CrossEntropyLoss()
            c = _apts_cfg(CfgNode)
        elif config == "cnn_apts_d":
            from dd4ml.models.cnn.simple_cnn import SimpleCNN
            mc = SimpleCNN.get_default_config()
            mc.input_channels, mc.input_height, mc.input_width, mc.output_classes = 3, 32, 32, 10
            model = SimpleCNN(mc)
            X, Y = W.cifar_shape_batch(per_rank_batch, 1234 + rank)
            crit = nn.CrossEntropyLoss()
            c = _apts_cfg(CfgNode)
        else:
            from dd4ml.models.gpt.nanogpt.model import GPT
            from dd4ml.utility.ml_utils import cross_entropy_transformers
            mc = GPT.get_default_config()
            mc.model_type, mc.vocab_size, mc.block_size = "gpt-mini", 65, 128
            model = GPT(mc)
            X, Y = W.token_batch(per_rank_batch, 1234 + rank)
            crit = cross_entropy_transformers
            c = _apts_cfg(CfgNode)
        c = APTS_D.setup_APTS_hparams(c)
        opt = APTS_D(model.parameters(), model=model, criterion=crit, device="cpu", nr_models=ws, glob_opt=c.glob_opt,
                     glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt, loc_opt_hparams=c.loc_opt_hparams,
                     glob_pass=True, foc=True, norm_type=2, max_loc_iters=3, max_glob_iters=1, tol=1e-6,
                     **c.apts_params)
        return (lambda: opt.step(X, Y)), opt
    if config == "pinn_apts_p":
        from dd4ml.datasets.pinn_allencahn import AllenCahn1DDataset
        from dd4ml.models.ffnn.pinn_ffnn import PINNFFNN
        from dd4ml.optimizers.apts_p import APTS_P
        from dd4ml.utility.pinn_allencahn_loss import AllenCahnPINNLoss
        ds = AllenCahn1DDataset(AllenCahn1DDataset.get_default_config())
        model = PINNFFNN(PINNFFNN.get_default_config())
        crit = AllenCahnPINNLoss()
        c = _apts_cfg(CfgNode, min_delta=0.01, max_delta=0.5, norm_type=math.inf, glob_second_order=False,
                      glob_dogleg=False, loc_second_order=False, loc_dogleg=False, mem_length=5, paper_tr_update=True)
        c = APTS_P.setup_APTS_hparams(c)
        opt = APTS_P(model.parameters(), model=model, criterion=crit, device="cpu", nr_models=ws, glob_opt=c.glob_opt,
                     glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt, loc_opt_hparams=c.loc_opt_hparams,
                     glob_pass=True, norm_type=math.inf, max_loc_iters=3, max_glob_iters=1, tol=1e-6, **c.apts_params)
        model = model.double()                                      # trainer.py:1203-1211
        opt.loc_model = opt.loc_model.double()
        x, y = ds.data.double(), ds.boundary_mask.double()
        x.requires_grad_(True)
        crit.current_x = x
        return (lambda: opt.step(inputs=x, labels=y)), opt
    if config == "deeponet_tr":
        from dd4ml.models.deeponet.deeponet import DeepONet
        from dd4ml.optimizers.tr import TR
        from dd4ml.utility import get_tr_hparams
        mc = DeepONet.get_default_config()
        mc.branch_input_dim = 50
        model = DeepONet(mc)
        b, t, y = W.deeponet_sine_batch()
        crit = nn.MSELoss()
        hp = get_tr_hparams(CfgNode(delta=0.1, max_delta=2.0, min_delta=1e-3, norm_type=2,
                                    glob_second_order=bool(overrides.get("second_order", True)), glob_dogleg=False,
                                    tol=1e-6))
        opt = TR(model.parameters(), **hp)

        def closure(compute_grad=True):
            opt.zero_grad()
            loss = crit(model((b, t)), y)
            if compute_grad:
                loss.backward()
            return loss
        return (lambda: opt.step(closure)[0]), opt
    raise ValueError(config)


def _worker(rank, ws, port, config, per_rank_batch, overrides, steps, warmup, threads, q):
    os.dup2(2, 1)                       # the reference prints banners (dprint); bench.py's stdout is one JSON line
    import torch
    import torch.distributed as dist
    torch.set_num_threads(max(1, threads))
    need_group = ws > 1 or config == "pinn_apts_p"      # mark_trainable calls dist.get_rank() unconditionally
    if need_group:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    step, opt = build(config, rank, ws, per_rank_batch, overrides)
    losses = []
    for _ in range(max(warmup, 1)):
        losses.append(float(step()))
    if need_group:
        dist.barrier()
    t0 = time.perf_counter()
    evals = 0.0
    for _ in range(steps):
        losses.append(float(step()))
        evals += float(getattr(opt, "grad_evals", 0.0))
    dt = time.perf_counter() - t0
    if need_group:
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        q.put(dict(seconds=dt, first_loss=losses[0], final_loss=losses[-1], grad_evals_per_step=evals / max(steps, 1)))


def run(config: str, n: int, steps: int, warmup: int, per_rank_batch: int, overrides: dict | None = None):
    """Times `steps` units of the metric on the unmodified reference with `n` gloo subdomain processes."""
    import socket

    import torch.multiprocessing as mp
    cores = os.cpu_count() or 1
    threads = max(1, cores // n)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    # Under torchrun the children must not inherit the launcher's rendezvous: with TORCHELASTIC_USE_AGENT_STORE
    # set, rank 0 of the children's own gloo group would not start its TCPStore (it expects the elastic agent to
    # host it) and every child waits forever.
    launcher_keys = [k for k in os.environ
                     if k.startswith("TORCHELASTIC_") or k in (
                         "RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "GROUP_RANK", "ROLE_RANK",
                         "ROLE_NAME", "LOCAL_WORLD_SIZE", "GROUP_WORLD_SIZE", "ROLE_WORLD_SIZE", "NODE_RANK",
                         "TORCH_NCCL_ASYNC_ERROR_HANDLING")]
    env_keep = {k: os.environ.pop(k) for k in launcher_keys}
    env_keep["OMP_NUM_THREADS"] = os.environ.get("OMP_NUM_THREADS", "")
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["CUDA_VISIBLE_DEVICES"] = ""            # the reference arm is the CPU path
    try:
        procs = [ctx.Process(target=_worker, args=(r, n, port, config, per_rank_batch, overrides or {}, steps, warmup,
                                                   threads, q)) for r in range(n)]
        for p in procs:
            p.start()
        out, waited = None, 0.0
        while out is None:                      # a child that died must not leave the parent waiting
            try:
                out = q.get(timeout=5.0)
            except Exception:  # noqa: BLE001  (queue.Empty)
                waited += 5.0
                dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
                if dead or waited > 3600:
                    for p in procs:
                        if p.is_alive():
                            p.terminate()
                    raise RuntimeError(f"reference arm: worker exit codes {dead} after {waited:.0f} s")
        for p in procs:
            p.join(timeout=120)
    finally:
        omp = env_keep.pop("OMP_NUM_THREADS")
        os.environ.update(env_keep)
        if omp:
            os.environ["OMP_NUM_THREADS"] = omp
        else:
            os.environ.pop("OMP_NUM_THREADS", None)
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    out.update(cores=threads * n, threads_per_process=threads, processes=n)
    return out
