#!/usr/bin/env python
"""Benchmark of the DD4ML trust-region hot path on B200 (contract: see the task statement).

    python bench.py --gpus 1 --steps K --warmup W                 # our arm, BASELINE configs[1]
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # N data subdomains, one per GPU
    python bench.py --impl reference ...                          # the unmodified reference on the host cores
    python bench.py --config {deeponet_tr,ffnn_apts_d,cnn_apts_d,gpt_apts_d,pinn_apts_p}
    python bench.py --impl aten                                   # diagnostic: the ATen restatement on cuda:0

Default workload (BASELINE.json configs[1], tests/config_files/config_apts_d.yaml of the reference):
MediumFFNN 784-[128]x8-10 (n = 217 354) on MNIST-shape synthetic data, APTS_D with
glob_opt = loc_opt = LSSR1_TR, second order + dogleg, mem_length 3, max_loc_iters 3,
max_glob_iters 1, foc, glob_pass, delta 0.1 in [0.01, 1.0]; one data subdomain of
PER_RANK_BATCH samples per GPU (weak scaling).  A "step" is one APTS_D outer iteration.
metric value = (#subdomains x outer iterations) / second, i.e. outer iterations per second at N=1.
The other four BASELINE configs are measured at their native sizes by `--config`, and in short form
inside the default N = 1 line (`other_configs`).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "APTS_D outer iters/sec"
UNIT = "subdomain outer iters/s"
L2_FLUSH_BYTES = 256 << 20

# name -> (per-rank batch, metric, unit, description)
CONFIGS = {
    "ffnn_apts_d": dict(batch=5000, metric=METRIC, unit=UNIT, k=3, workload=(
        "BASELINE configs[1]: MediumFFNN 784-[128]x8-10 (n=217354) on MNIST-shape synthetic data, APTS_D "
        "glob=loc=LSSR1_TR second-order+dogleg, mem_length 3, max_loc_iters 3, max_glob_iters 1, foc, glob_pass")),
    "cnn_apts_d": dict(batch=200, metric=METRIC, unit=UNIT, k=3, workload=(
        "BASELINE configs[2]: SimpleCNN (n=620810) on CIFAR-10-shape synthetic data, batch 200 per subdomain, APTS_D "
        "glob=loc=LSSR1_TR second-order+dogleg, mem_length 3, max_loc_iters 3, max_glob_iters 1, foc, glob_pass; "
        "fp32 convolutions (TF32 off), BatchNorm + Dropout(0.5) active")),
    "gpt_apts_d": dict(batch=16, metric=METRIC, unit=UNIT, k=3, workload=(
        "BASELINE configs[3]: minGPT gpt-mini 6L/6H/192 (n=2719104), 16 x 128 synthetic tokens per subdomain, "
        "dropout 0.1 active, APTS_D glob=loc=LSSR1_TR second-order+dogleg, mem_length 3, max_loc_iters 3, "
        "max_glob_iters 1, foc, glob_pass")),
    "pinn_apts_p": dict(batch=1002, metric="APTS_P outer iters/sec", unit="outer iters/s", k=0, workload=(
        "BASELINE configs[4]: PINN Allen-Cahn, PINNFFNN 1-[20]x8-1 (n=3001), 1000 interior + 2 boundary points, APTS_P "
        "(config_apts_p.yaml: glob=loc=LSSR1_TR first-order, paper radius update, inf-norm, delta 0.1 in [0.01,0.5]), "
        "built in fp32 and converted to fp64 like the reference Trainer does")),
    "deeponet_tr": dict(batch=1000, metric="TR steps/sec", unit="TR steps/s", k=5, workload=(
        "BASELINE configs[0]: DeepONet (n=11712) on the deeponet_sine synthetic operator data, batch 1000, plain TR "
        "second-order (L-SR1 mem 5, OBS), delta 0.1 in [1e-3, 2], float targets + MSE")),
}
PER_RANK_BATCH = CONFIGS["ffnn_apts_d"]["batch"]


def env_int(name, default):
    return int(os.environ.get(name, default))


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------- our arm: jobs
class Job:
    """One BASELINE config on one rank: `step(batch) -> loss` is the call a user makes (the public
    optimizer API), `host_batch` the pinned host tensors of this rank's (fixed) batch."""

    def __init__(self, name, model, opt, host_batch, step, upload=None, states=()):
        self.name, self.model, self.opt, self.host_batch, self.step = name, model, opt, host_batch, step
        self.upload = upload
        self.states = list(states)
        self.n = sum(p.numel() for p in model.parameters())

    def sync_count(self):
        return sum(s.syncs for s in self.states)


def _apts_d(model, crit, device, world, **over):
    from dd4ml_b200 import APTS_D
    from dd4ml_b200.workloads import apts_d_yaml_config
    c = APTS_D.setup_APTS_hparams(apts_d_yaml_config(**over))
    return APTS_D(model.parameters(), model=model, criterion=crit, device=device, nr_models=world,
                  glob_opt=c.glob_opt, glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt,
                  loc_opt_hparams=c.loc_opt_hparams, glob_pass=c.glob_pass, foc=c.foc, norm_type=c.norm_type,
                  max_loc_iters=c.max_loc_iters, max_glob_iters=c.max_glob_iters, tol=c.tol, **c.apts_params)


def build_optimizer(device, world, rank=None):
    """(model, optimizer) of the default bench workload: what the diagnostics under tools/ drive."""
    job = build_job("ffnn_apts_d", device, world, env_int("RANK", 0) if rank is None else rank)
    return job.model, job.opt


def build_job(name, device, world, rank, labels="random", batch=None):
    import torch.nn as nn
    from dd4ml_b200 import workloads as W
    spec = CONFIGS[name]
    B = batch or spec["batch"]
    torch.manual_seed(42)                                   # same initial weights on every rank
    pin = lambda *ts: tuple(t.pin_memory() for t in ts)  # noqa: E731
    dev = lambda ts: tuple(t.to(device, non_blocking=True) for t in ts)  # noqa: E731
    if name == "ffnn_apts_d":
        model = W.MediumFFNN().to(device)
        X, Y = W.mnist_shape_batch(B, 1234 + rank)
        if labels == "teacher":
            Y = W.teacher_labels(X)
        opt = _apts_d(model, nn.CrossEntropyLoss(), device, world)
        job = Job(name, model, opt, pin(X, Y), lambda b: opt.step(b[0], b[1]), dev)
    elif name == "cnn_apts_d":
        model = W.SimpleCNN().to(device)
        opt = _apts_d(model, nn.CrossEntropyLoss(), device, world)
        job = Job(name, model, opt, pin(*W.cifar_shape_batch(B, 1234 + rank)), lambda b: opt.step(b[0], b[1]), dev)
    elif name == "gpt_apts_d":
        model = W.MiniGPT().to(device)
        opt = _apts_d(model, W.token_cross_entropy, device, world)
        job = Job(name, model, opt, pin(*W.token_batch(B, 1234 + rank)), lambda b: opt.step(b[0], b[1]), dev)
    elif name == "pinn_apts_p":
        from dd4ml_b200 import APTS_P
        model = W.PINNNet().to(device)
        crit = W.AllenCahnLoss()
        c = APTS_P.setup_APTS_hparams(W.Cfg(
            delta=0.1, min_delta=0.01, max_delta=0.5, norm_type=math.inf, glob_second_order=False, glob_dogleg=False,
            loc_second_order=False, loc_dogleg=False, mem_length=5, max_wolfe_iters=5, max_zoom_iters=5, tol=1e-6,
            paper_tr_update=True, glob_opt="lssr1_tr", loc_opt="lssr1_tr"))
        opt = APTS_P(model.parameters(), model=model, criterion=crit, device=device, nr_models=world,
                     glob_opt=c.glob_opt, glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt,
                     loc_opt_hparams=c.loc_opt_hparams, glob_pass=True, norm_type=math.inf, max_loc_iters=3,
                     max_glob_iters=1, tol=1e-6, **c.apts_params)
        model = model.double()                               # trainer.py:1203-1211
        opt.loc_model = opt.loc_model.double()
        x, flag = W.allen_cahn_points()

        def upload(b):
            xd = b[0].to(device, non_blocking=True).double().requires_grad_(True)   # trainer.py:987-989
            crit.current_x = xd
            return xd, b[1].to(device, non_blocking=True).double()
        def pinn_step(b):
            crit.current_x = b[0]                            # trainer.py:997-998: set for every batch
            return opt.step(inputs=b[0], labels=b[1])
        job = Job(name, model, opt, pin(x, flag), pinn_step, upload)
    elif name == "deeponet_tr":
        from dd4ml_b200 import TR
        from dd4ml_b200.utility.optimizer_utils import get_tr_hparams
        model = W.DeepONet().to(device)
        crit = nn.MSELoss()
        hp = get_tr_hparams(W.Cfg(delta=0.1, max_delta=2.0, min_delta=1e-3, norm_type=2, glob_second_order=True,
                                  glob_dogleg=False, tol=1e-6))
        opt = TR(model.parameters(), **hp)

        def step(b):
            def closure(compute_grad=True):
                opt.zero_grad()
                loss = crit(model((b[0], b[1])), b[2])
                if compute_grad:
                    loss.backward()
                return loss
            return opt.step(closure)[0]
        job = Job(name, model, opt, pin(*W.deeponet_sine_batch(B)), step, dev)
    else:
        raise SystemExit(f"unknown --config {name}")
    opt_states = [getattr(opt, "_ctl", None), getattr(getattr(opt, "glob_opt", None), "_ds", None),
                  getattr(getattr(opt, "loc_opt", None), "_ds", None), getattr(opt, "_ds", None)]
    job.states = [s for s in opt_states if s is not None]
    return job


def timed_region(fn, steps, world, device, flush):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        flush.zero_()                            # L2 flush: 256 MiB memset before every step
        fn(i)
    e1.record()
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        dist.barrier()
    return ms, wall


def algorithmic_bytes(name, n, k, b=4):
    """Algorithmic HBM bytes per launch: vectors read + written, k live pairs."""
    return {
        "dd4ml_tr_propose": (2 * k + 1) * b * n,
        "dd4ml_tr_apply": (2 * k + 2) * b * n,          # direction only (LSSR1_TR); +2 with w += p
        "dd4ml_tr_decide": (2 * k + 6) * b * n,
        "dd4ml_lssr1_begin": (2 * k + 8) * b * n,
        "dd4ml_lssr1_trial": 3 * b * n,
        "dd4ml_lssr1_eval": 3 * b * n,
        "dd4ml_apts_residual": 5 * b * n,
        "dd4ml_apts_foc": 3 * b * n,
        "dd4ml_apts_pack": 3 * b * n,
        "dd4ml_apts_trial": 3 * b * n,
        "dd4ml_apts_control": 2 * b * n,
    }.get(name)


def large_n_roofline(device, kernel, peak_gbs, peak_src, n=1 << 26, k=3, iters=20):
    """The same fused kernels on vectors far larger than L2 (HBM-bound regime), timed with CUDA
    events on the launching stream; inputs (>= 10 x 256 MiB) exceed the 126 MB L2."""
    from dd4ml_b200 import _lib
    from dd4ml_b200.optimizers.lsr1 import LSR1
    lib = _lib.get()
    h = LSR1(gamma=1e-3, memory_length=k, tol=1e-6, device=device)
    g = torch.Generator(device=device).manual_seed(1)
    for _ in range(k):
        s = torch.randn(n, device=device, generator=g) * 0.01
        y = s + 0.1 * torch.randn(n, device=device, generator=g) * 0.01
        h.update_memory(s, y)
    assert len(h._S) == k
    ds = h.ds
    ds.set(delta=0.1, min_delta=0.01, max_delta=1.0, tol=1e-6, second_order=1.0, dogleg=1.0)
    r = lambda: torch.randn(n, device=device, generator=g)  # noqa: E731
    w, old_w, gv, prev_g, p = r(), r(), r(), r(), torch.empty(n, device=device)
    loss = torch.tensor(1.0, device=device)
    st = _lib.stream()
    calls = {
        "dd4ml_lssr1_begin": lambda: lib.lssr1_begin(ds.sp, ds.wp, w.data_ptr(), old_w.data_ptr(), old_w.data_ptr(), gv.data_ptr(), prev_g.data_ptr(),
                                                     prev_g.data_ptr(), 1, h.S_ptr, h.Y_ptr, loss.data_ptr(), 0, n,
                                                     h.ld, 0, 0, 0, k, -1.0, 0.0, 0.0, st),
        "dd4ml_tr_propose": lambda: lib.tr_propose(ds.sp, ds.wp, gv.data_ptr(), h.S_ptr, h.Y_ptr, n, h.ld, 0, 0, 1, k, st),
        "dd4ml_tr_apply": lambda: lib.tr_apply(ds.sp, ds.wp, gv.data_ptr(), h.S_ptr, h.Y_ptr, p.data_ptr(), None, n,
                                               h.ld, 0, 0, 0, k, st),
        "dd4ml_lssr1_trial": lambda: lib.lssr1_trial(w.data_ptr(), old_w.data_ptr(), p.data_ptr(), 0.5, n, 0, 0, st),
        "dd4ml_lssr1_eval": lambda: lib.lssr1_eval(ds.sp, ds.wp, gv.data_ptr(), p.data_ptr(), prev_g.data_ptr(),
                                                   loss.data_ptr(), 0, n, 0, st),
    }
    trial_l = torch.tensor(0.5, device=device)

    def decide():
        ds.set(pred=1.0, converged=0.0)          # rho = (1.0 - 0.5) / 1.0 -> accepted: the heavy path
        lib.tr_decide(ds.sp, ds.wp, loss.data_ptr(), trial_l.data_ptr(), 0, gv.data_ptr(), prev_g.data_ptr(),
                      old_w.data_ptr(), p.data_ptr(), w.data_ptr(), h.S_ptr, h.Y_ptr, None, n, h.ld, 0, 0, 0, k, st)
    calls["dd4ml_tr_decide"] = decide
    out = {}
    two_kernel = ("dd4ml_lssr1_begin", "dd4ml_tr_propose", "dd4ml_tr_decide")

    def time_it(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / iters

    for name, fn in calls.items():
        if name == "dd4ml_tr_apply":
            # the direction pass streams the pairs whose coefficients are non-zero (a Cauchy-point or first-order
            # step streams none): give every live pair a coefficient so that the launch reads what
            # `algorithmic_bytes` counts, whatever the solve on the synthetic memory came out as
            lib.tr_propose(ds.sp, ds.wp, gv.data_ptr(), h.S_ptr, h.Y_ptr, n, h.ld, 0, 0, 1, k, st)
            full = ds.read(full=True)
            coef = [0.0] * len(full.cS)
            for sl in (int(o) for o in full.order[:int(full.k)]):
                coef[sl] = 1e-3
            ds.set_array("cS", coef)
            ds.set_array("cY", coef)
            ds.set(cg=-1.0, converged=0.0)
        ms_entry = time_it(fn)
        kk = int(ds.read().k)
        ms = ms_entry
        if name in two_kernel:
            # these entry points are TWO kernels: the HBM-bound streaming reduction and the one-CTA,
            # latency-bound k-space kernel behind it.  The roofline fraction is the streaming kernel's
            # (timed alone: k-space kernel switched off through the library's measurement hook); the
            # whole entry point is reported next to it.
            state_copy = ds.state.clone()
            lib.cdll.dd4ml_debug_kspace(0)
            try:
                ms = time_it(fn)
            finally:
                lib.cdll.dd4ml_debug_kspace(1)
                ds.state.copy_(state_copy)
        byt = algorithmic_bytes(name, n, kk)
        out[name] = {"n": n, "k": kk, "ms": ms, "bytes": byt, "achieved": byt / ms / 1e6,
                     "frac": byt / ms / 1e6 / peak_gbs, "ms_entry_with_kspace_kernel": ms_entry,
                     "frac_entry_with_kspace_kernel": byt / ms_entry / 1e6 / peak_gbs,
                     "kspace_kernel_us": 1e3 * (ms_entry - ms)}
    sel = out.get(kernel) or out["dd4ml_lssr1_begin"]
    # DRAM traffic of the same kernel at the same (n, k) from the committed `ncu --set full` capture
    traffic = None
    try:
        key = {"dd4ml_lssr1_begin": "LssrBeginOp", "dd4ml_tr_propose": "ProposeOp", "dd4ml_tr_apply": "ApplyOp",
               "dd4ml_lssr1_eval": "EvalOp", "dd4ml_lssr1_trial": "TrialOp"}[kernel if kernel in out else "dd4ml_lssr1_begin"]
        for fname in ("ncu_full_r2_summary.json", "ncu_full_r1_summary.json"):
            path = os.path.join(ROOT, "profiles", fname)
            if not os.path.exists(path):
                continue
            summ = json.load(open(path))[key]
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tr = 0.0
            for f in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = summ[f].split()
                tr += float(v) * unit[u]
            if n == 1 << 26 and k == 3:
                traffic = tr
            break
    except Exception:  # noqa: BLE001
        traffic = None
    roof = {"bound": "hbm", "achieved": sel["achieved"], "peak": peak_gbs, "unit": "GB/s",
            "frac": sel["frac"], "traffic": traffic, "kernel": kernel if kernel in out else "dd4ml_lssr1_begin",
            "workload": f"fp32 n=2^26 k={sel['k']} (inputs >> L2)", "peak_source": peak_src,
            "bytes_per_launch": sel["bytes"], "ms_per_launch": sel["ms"],
            "what": "the streaming reduction kernel of the entry point, timed alone with CUDA events",
            "entry_point_with_kspace_kernel": {"ms": sel["ms_entry_with_kspace_kernel"],
                                               "frac": sel["frac_entry_with_kspace_kernel"],
                                               "kspace_kernel_us": sel["kspace_kernel_us"]}}
    return roof, out


# ---------------------------------------------------------------------------------- measurement of one job
def measure(job, args, world, device, flush, rank, local, full=True, make_job=None):
    """Warm-up, the timed region (inputs resident), the end-to-end region (host batches), and one
    profiled pass (CUDA events around every own launch, every closure, every collective)."""
    import torch.distributed as dist
    from dd4ml_b200 import _lib
    from dd4ml_b200.graphs import GraphedEval
    from dd4ml_b200.optimizers import apts_base
    lib = _lib.get()
    opt = job.opt
    batch = job.upload(job.host_batch)
    torch.cuda.synchronize(device)
    losses, evals, loc_evals = [], [], []

    def step_resident(i):
        losses.append(job.step(batch))
        evals.append(float(getattr(opt, "grad_evals", 0.0)))
        loc_evals.append(float(getattr(opt, "loc_grad_evals", 0.0)))     # THIS rank's local evaluations

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0 and full:
        sampler.start()
    l0, s0 = lib.launches, job.sync_count()
    del evals[:], loc_evals[:]
    ms, wall = timed_region(step_resident, args.steps, world, device, flush)
    evals_per_step = sum(evals) / max(len(evals), 1)
    loc_evals_per_step = sum(loc_evals) / max(len(loc_evals), 1)
    launches = lib.launches - l0
    host_syncs = (job.sync_count() - s0) / args.steps
    clocks = sampler.stop() if (rank == 0 and full) else None
    spec = CONFIGS[job.name]
    value = world * args.steps / (ms / 1e3)

    # ---- end-to-end: host batch -> device every step, loss read back every step (the call a user makes).
    # The work per outer iteration depends on where the run is on its trajectory (line-search evaluations),
    # so the end-to-end region runs on a FRESH job with the same seeds: same warm-up, same outer iterations
    # as the resident region above.
    ejob = make_job() if make_job is not None else job
    host_loss = torch.empty((), dtype=torch.float64).pin_memory()
    feed = None
    if len(ejob.host_batch) == 2 and ejob.name != "pinn_apts_p":
        # the public feed: pinned host batches through HostBatchPrefetcher -- the copy of the next step's
        # batch runs on a side stream under the current step; one H2D copy per step
        import itertools
        from dd4ml_b200.dataloaders import HostBatchPrefetcher
        feed = iter(HostBatchPrefetcher(itertools.repeat(ejob.host_batch), device))
        nxt = lambda: next(feed)  # noqa: E731
    else:
        nxt = lambda: ejob.upload(ejob.host_batch)  # noqa: E731

    def step_e2e(i):
        loss = ejob.step(nxt())
        host_loss.copy_(loss.detach().double(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for i in range(args.warmup if ejob is not job else 2):
        step_e2e(i)
    ms_e2e, _ = timed_region(step_e2e, args.steps, world, device, flush)
    e2e = {"value": world * args.steps / (ms_e2e / 1e3), "unit": spec["unit"],
           "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in ejob.host_batch),
           "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps,
           "iterations": "the same outer iterations as `value` (fresh job, same seeds and warm-up)"
                         if ejob is not job else "the outer iterations after the resident region"}
    if ejob is not job:
        del ejob, feed, nxt
        torch.cuda.empty_cache()

    # ---- profiled pass: per-kernel device times inside the step, closures, collectives
    lib.prof, GraphedEval.PROF, apts_base.COMM_PROF = {}, [], []
    nprof = min(args.steps, 10)
    ev_step = []
    for i in range(nprof):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_resident(i)
        b.record()
        ev_step.append((a, b))
    torch.cuda.synchronize(device)
    prof, cl, comm = lib.prof, GraphedEval.PROF, apts_base.COMM_PROF
    lib.prof, GraphedEval.PROF, apts_base.COMM_PROF = None, None, None
    n, kmem = job.n, spec["k"]
    esz = 8 if next(job.model.parameters()).dtype == torch.float64 else 4
    native = {}
    for name, evs in prof.items():
        ts = [a.elapsed_time(b) for a, b in evs]
        byt = algorithmic_bytes(name, n, kmem, esz)
        native[name] = {"launches_per_step": len(ts) / nprof, "avg_us": 1e3 * sum(ts) / len(ts),
                        "min_us": 1e3 * min(ts), "median_us": 1e3 * statistics.median(ts), "max_us": 1e3 * max(ts),
                        "total_ms_per_step": sum(ts) / nprof,
                        "achieved_gbs": None if not byt else byt / (sum(ts) / len(ts)) / 1e6}
    prof_step_ms = sum(a.elapsed_time(b) for a, b in ev_step) / nprof
    closure_ms = sum(a.elapsed_time(b) for a, b in cl) / nprof if cl else None
    comm_ts = [a.elapsed_time(b) for a, b, k in comm if k > 0]          # NCCL all-reduces
    comb_ts = [a.elapsed_time(b) for a, b, k in comm if k < 0]          # fused peer-memory combine (+ its barrier)
    own_ms = sum(v["total_ms_per_step"] for v in native.values())
    dominant = max(native, key=lambda k: native[k]["total_ms_per_step"]) if native else "dd4ml_lssr1_begin"
    # per-rank view (N > 1): the subdomains' local phases take data-dependent numbers of evaluations and
    # every collective waits for the slowest rank -- the spread below is what the weak-scaling curve pays
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "local_evals_per_step": loc_evals_per_step,
                "closure_ms_per_step": closure_ms, "closures_per_step": (len(cl) / nprof) if cl else None,
                "exchange_ms_per_step": (sum(comm_ts) + sum(comb_ts)) / nprof,
                "own_kernel_ms_per_step": own_ms}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        le = [r["local_evals_per_step"] for r in allr]
        ex = [r["exchange_ms_per_step"] for r in allr]
        per_rank = {"local_evals_per_step": {"min": min(le), "mean": sum(le) / world, "max": max(le)},
                    "exchange_ms_per_step_incl_waiting": {"min": min(ex), "mean": sum(ex) / world, "max": max(ex)},
                    "ranks": allr}
    res = {
        "metric": spec["metric"], "value": value, "unit": spec["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "e2e": e2e, "gpu_launches": launches,
        "host_syncs_per_step": host_syncs, "grad_evals_per_step": evals_per_step,
        "ms_per_grad_eval": (ms / args.steps) / evals_per_step if evals_per_step else None,
        # device time of the untouched PyTorch forward/backward closures vs everything else in the step
        "closure_ms_per_step": closure_ms, "closures_per_step": (len(cl) / nprof) if cl else None,
        "optimizer_side_ms_per_step": None if closure_ms is None else prof_step_ms - closure_ms,
        "profiled_ms_per_step": prof_step_ms,
        "allreduce": None if not comm_ts else {
            "per_step": len(comm_ts) / nprof, "avg_us": 1e3 * sum(comm_ts) / len(comm_ts),
            "median_us": 1e3 * statistics.median(comm_ts), "max_us": 1e3 * max(comm_ts),
            "total_ms_per_step": sum(comm_ts) / nprof, "payload_elems": max(c[2] for c in comm)},
        "p2p_combine": None if not comb_ts else {
            "per_step": len(comb_ts) / nprof, "avg_us": 1e3 * sum(comb_ts) / len(comb_ts),
            "median_us": 1e3 * statistics.median(comb_ts), "max_us": 1e3 * max(comb_ts),
            "what": "device barrier + ONE kernel: sum of all subdomains' packed steps over NVLink peer memory, "
                    "applied to w0 (replaces pack -> ncclAllReduce -> apts_trial)"},
        "exchange": getattr(opt, "_p2p_note", None),
        "wall_ms_per_step": wall * 1e3 / args.steps, "per_rank": per_rank,
        "first_loss": float(losses[0]), "final_loss": float(losses[-1]),
        "roofline_native": {"n": n, "k": kmem, "dominant": dominant, "own_kernel_ms_per_step": own_ms,
                            "own_kernel_share_of_step": own_ms / prof_step_ms, "kernels": native},
    }
    return res, clocks, dominant


def reference_line(args, config, n, overrides=None):
    """The reference arm's numbers: unmodified reference from baseline/_ref when present (kind
    "reference"), else the oracle port (kind "port")."""
    from tools import bench_reference as R
    spec = CONFIGS[config]
    threads = os.cpu_count() or 1
    if R.available():
        out = R.run(config, n, args.steps, max(args.warmup, 1), spec["batch"], overrides)
        value = n * args.steps / out["seconds"]
        kind, cores = "reference", out["cores"]
        sample = (f"{args.steps} units of the full workload after {max(args.warmup, 1)} warm-up, unmodified "
                  f"dd4ml (baseline/_ref), {out['processes']} gloo process(es) x {out['threads_per_process']} threads")
        extra = {k: out[k] for k in ("first_loss", "final_loss", "grad_evals_per_step")}
        ms = out["seconds"] / args.steps * 1e3
    else:
        if config != "ffnn_apts_d":
            return None
        opt, fgs = port_baseline(n_ranks=n, threads=threads)
        for _ in range(max(args.warmup, 1)):
            opt.step(fgs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss = opt.step(fgs)
        dt = time.perf_counter() - t0
        value, kind, cores, ms = n * args.steps / dt, "port", threads, dt / args.steps * 1e3
        sample = f"{args.steps} outer iterations of the full workload ({n} simulated subdomain(s) in one process)"
        extra = {"final_loss": float(loss)}
    return dict(value=value, kind=kind, cores=cores, sample=sample, ms_per_step=ms, **extra)


def port_baseline(n_ranks=1, threads=None):
    """The oracle port of the reference's CPU path (fallback when baseline/_ref is absent)."""
    import torch.nn as nn
    import oracle
    from dd4ml_b200.workloads import MediumFFNN, mnist_shape_batch
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(42)
    models = [MediumFFNN() for _ in range(n_ranks)]
    w0 = torch.cat([p.detach().reshape(-1) for p in models[0].parameters()]).clone()
    crit = nn.CrossEntropyLoss()

    def make_fg(model, X, Y):
        ps = list(model.parameters())

        def fg(w):
            off = 0
            with torch.no_grad():
                for p in ps:
                    k = p.numel()
                    p.copy_(w[off:off + k].view_as(p)); off += k
            for p in ps:
                p.grad = None
            loss = crit(model(X), Y)
            loss.backward()
            return loss.detach(), torch.cat([p.grad.reshape(-1) for p in ps])
        return fg

    fgs = [make_fg(models[r], *mnist_shape_batch(CONFIGS["ffnn_apts_d"]["batch"], 1234 + r)) for r in range(n_ranks)]
    ghp = oracle.lssr1_hparams(0.1, 0.01, 1.0, second_order=True, dogleg=True, mem_length=3)
    lhp = oracle.lssr1_loc_hparams(0.1, 0.01, 1.0, n_ranks, second_order=True, dogleg=True, mem_length=3)
    opt = oracle.APTSDOracle(w0, n_ranks, "lssr1_tr", ghp, "lssr1_tr", lhp, **oracle.apts_hparams(0.1, 0.01, 1.0),
                             glob_pass=True, foc=True, norm_type=2, max_loc_iters=3, max_glob_iters=1, tol=1e-6)
    return opt, fgs


def workload_config(config, n):
    spec = CONFIGS[config]
    return {"workload": spec["workload"], "name": config, "subdomains": n, "per_rank_batch": spec["batch"],
            "parallelism": f"{config.split('_', 1)[1] if '_' in config else config} x{n} (one subdomain per GPU)",
            "value_definition": "subdomains x units / s (= units / s at 1 GPU)",
            "l2_flush": "256 MiB memset before every step, inside the timed region"}


def run_reference(args):
    """`--impl reference`: rank 0 alone runs it; under torchrun the other ranks exit 0."""
    if env_int("RANK", 0) != 0:
        return
    n = args.gpus
    r = reference_line(args, args.config, n, {"labels": args.labels})
    spec = CONFIGS[args.config]
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref missing and no port for this config"}))
        return
    print(json.dumps({
        "impl": "reference", "metric": spec["metric"], "value": r["value"], "unit": spec["unit"], "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.config == "pinn_apts_p" else "f32",
        "data": "synthetic", "config": dict(workload_config(args.config, n), labels=args.labels),
        "cpu_baseline": {"value": r["value"], "unit": spec["unit"], "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": spec["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_loss": r.get("final_loss"), "first_loss": r.get("first_loss"),
        "grad_evals_per_step": r.get("grad_evals_per_step")}))


def run_aten(args):
    """`--impl aten` (diagnostic, N = 1): the ATen restatement of the path (the oracle port: the
    reference's op sequence over flat vectors) with every vector and the model on cuda:0 -- the same
    GPU doing the forward/backward, PyTorch's stock kernels doing the optimizer math.  What the
    dd4ml_b200 kernels buy is `value` of the default arm over this line's."""
    import torch.nn as nn
    import oracle
    from dd4ml_b200.workloads import MediumFFNN, mnist_shape_batch
    if args.config != "ffnn_apts_d":
        raise SystemExit("--impl aten: ffnn_apts_d only")
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = MediumFFNN().to(dev)
    ps = list(model.parameters())
    w0 = torch.cat([p.detach().reshape(-1) for p in ps]).clone()
    X, Y = (t.to(dev) for t in mnist_shape_batch(CONFIGS["ffnn_apts_d"]["batch"], 1234))
    crit = nn.CrossEntropyLoss()

    def fg(w):
        torch.nn.utils.vector_to_parameters(w, ps)
        for p in ps:
            p.grad = None
        loss = crit(model(X), Y)
        loss.backward()
        return loss.detach(), torch.cat([p.grad.reshape(-1) for p in ps])
    ghp = oracle.lssr1_hparams(0.1, 0.01, 1.0, second_order=True, dogleg=True, mem_length=3)
    lhp = oracle.lssr1_loc_hparams(0.1, 0.01, 1.0, 1, second_order=True, dogleg=True, mem_length=3)
    opt = oracle.APTSDOracle(w0, 1, "lssr1_tr", ghp, "lssr1_tr", lhp, **oracle.apts_hparams(0.1, 0.01, 1.0),
                             glob_pass=True, foc=True, norm_type=2, max_loc_iters=3, max_glob_iters=1, tol=1e-6)
    for _ in range(max(args.warmup, 3)):
        opt.step([fg])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = opt.step([fg])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"impl": "aten", "metric": METRIC, "value": args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                      "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": workload_config("ffnn_apts_d", 1), "final_loss": float(loss),
                      "note": "same GPU, model + vectors on cuda:0, optimizer math = stock ATen kernels (oracle port)"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="ffnn_apts_d", choices=sorted(CONFIGS))
    ap.add_argument("--labels", default="random", choices=["random", "teacher"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other configs / teacher labels")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--teacher-extra", action="store_true",
                    help="also time the same workload with teacher labels (the N >= 2 stall at ln 10 is the same)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "aten":
        return run_aten(args)

    import torch.distributed as dist
    rank, local, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- dd4ml_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = False          # the path is fp32 (reference arithmetic); no TF32 convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from dd4ml_b200 import _lib
    _lib.get()

    # ---- multi-rank parity against goldens of the unmodified reference, before anything is timed
    parity = None
    if world > 1 and not args.no_parity:
        from tools import bench_parity
        parity, ok = bench_parity.run(rank, world, str(device))
        if not ok:
            if rank == 0:
                print(json.dumps({"parity_check": parity, "error": "multi-rank parity mismatch"}))
            dist.barrier()
            dist.destroy_process_group()
            raise SystemExit(3)

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
    job = build_job(args.config, device, world, rank, labels=args.labels)
    res, clocks, dominant = measure(job, args, world, device, flush, rank, local,
                                    make_job=lambda: build_job(args.config, device, world, rank, labels=args.labels))

    # ---- equal-work scaling diagnostic: the same workload with learnable (teacher) labels
    teacher = None
    if args.config == "ffnn_apts_d" and args.labels == "random" and args.teacher_extra:
        tjob = build_job("ffnn_apts_d", device, world, rank, labels="teacher")
        targs = argparse.Namespace(steps=min(args.steps, 20), warmup=3)
        tres, _, _ = measure(tjob, targs, world, device, flush, rank, local, full=False)
        teacher = {k: tres[k] for k in ("value", "unit", "ms_per_step", "grad_evals_per_step", "ms_per_grad_eval",
                                        "first_loss", "final_loss", "host_syncs_per_step", "allreduce")}
        teacher["labels"] = "argmax of a fixed random teacher MLP (learnable: no stall at ln 10 for N >= 2)"
        del tjob

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"

    # ---- the other BASELINE configs at their native sizes, short form (N = 1 only)
    others = None
    if world == 1 and args.config == "ffnn_apts_d" and not args.no_extra:
        others = {}
        for name in ("gpt_apts_d", "cnn_apts_d", "pinn_apts_p", "deeponet_tr"):
            try:
                oj = build_job(name, device, 1, 0)
                oargs = argparse.Namespace(steps=10, warmup=3)
                ores, _, odom = measure(oj, oargs, 1, device, flush, 0, local, full=False)
                kn = ores["roofline_native"]["kernels"]
                others[name] = {
                    "workload": CONFIGS[name]["workload"], "n": oj.n,
                    **{k: ores[k] for k in ("metric", "value", "unit", "ms_per_step", "grad_evals_per_step",
                                            "host_syncs_per_step", "closure_ms_per_step",
                                            "optimizer_side_ms_per_step", "first_loss", "final_loss")},
                    "e2e_value": ores["e2e"]["value"],
                    "own_kernel_ms_per_step": ores["roofline_native"]["own_kernel_ms_per_step"],
                    "dominant": odom,
                    "kernels": {k: {f: v[f] for f in ("launches_per_step", "median_us", "min_us", "achieved_gbs")}
                                for k, v in kn.items()},
                    # fraction of the measured HBM peak the dominant own kernel reaches INSIDE the step (after the
                    # L2 flush + closures: cold vectors at this config's native size)
                    "dominant_frac_of_hbm_peak": (None if not kn.get(odom, {}).get("achieved_gbs") else
                                                  kn[odom]["achieved_gbs"] / peak),
                }
                del oj
            except Exception as e:  # noqa: BLE001
                others[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    roofline, sweep, roof_gpt = (None, None, None)
    if not args.no_sweep:
        roofline, sweep = large_n_roofline(device, dominant, peak, peak_src)
        if world == 1:
            # the same measurement at the largest BASELINE vector length (gpt-mini, n = 2 719 104: 152 MB per
            # lssr1_begin launch, above the 126 MB L2): the HBM-bound regime on a real config's size
            rg, sg = large_n_roofline(device, dominant, peak, peak_src, n=2719104, k=3, iters=50)
            rg["workload"] = "fp32 n=2719104 (minGPT gpt-mini, BASELINE configs[3]) k=3, launches back to back"
            rg.pop("traffic", None)
            roof_gpt = {"roofline": rg, "kernels": {k: {"ms": v["ms"], "frac": v["frac"],
                                                        "frac_entry_with_kspace_kernel": v["frac_entry_with_kspace_kernel"]}
                                                    for k, v in sg.items()}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cargs = argparse.Namespace(steps=60 if args.config in ("ffnn_apts_d", "pinn_apts_p", "deeponet_tr") else 5,
                                   warmup=1)
        r = reference_line(cargs, args.config, 1, {"labels": args.labels})
        if r is not None:
            cpu = {"value": r["value"], "unit": CONFIGS[args.config]["unit"], "cores": r["cores"], "kind": r["kind"],
                   "sample": r["sample"]}
    out = {
        **{k: res[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step")},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.config == "pinn_apts_p" else "f32", "data": "synthetic",
        "config": dict(workload_config(args.config, world), labels=args.labels),
        "outer_iters_per_sec": args.steps / (res["ms_per_step"] * args.steps / 1e3),
        **{k: res[k] for k in ("wall_ms_per_step", "e2e", "gpu_launches", "host_syncs_per_step", "grad_evals_per_step",
                               "ms_per_grad_eval", "closure_ms_per_step", "closures_per_step",
                               "optimizer_side_ms_per_step", "profiled_ms_per_step", "allreduce", "p2p_combine",
                               "exchange", "per_rank")},
        "clocks": clocks, "roofline": roofline, "roofline_native": res["roofline_native"],
        "roofline_sweep": sweep, "roofline_gpt_mini_size": roof_gpt, "cpu_baseline": cpu, "parity_check": parity, "teacher_labels": teacher,
        "other_configs": others,
        "final_loss": res["final_loss"], "first_loss": res["first_loss"],
    }
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
