#!/usr/bin/env python
"""Benchmark of the DD4ML trust-region hot path on B200 (contract: see the task statement).

    python bench.py --gpus 1 --steps K --warmup W                 # our arm
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # N data subdomains, one per GPU
    python bench.py --impl reference ...                          # CPU baseline arm (oracle port)

Workload (BASELINE.json configs[1], tests/config_files/config_apts_d.yaml of the reference):
MediumFFNN 784-[128]x8-10 (n = 217 354) on MNIST-shape synthetic data, APTS_D with
glob_opt = loc_opt = LSSR1_TR, second order + dogleg, mem_length 3, max_loc_iters 3,
max_glob_iters 1, foc, glob_pass, delta 0.1 in [0.01, 1.0]; one data subdomain of
PER_RANK_BATCH samples per GPU (weak scaling).  A "step" is one APTS_D outer iteration.
metric value = (#subdomains x outer iterations) / second, i.e. outer iterations per second at N=1.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_RANK_BATCH = 5000          # config_apts_d.yaml: batch_size 10000 over num_subdomains 2
METRIC = "APTS_D outer iters/sec"
UNIT = "subdomain outer iters/s"
L2_FLUSH_BYTES = 256 << 20


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


# ---------------------------------------------------------------------------------- our arm
def build_optimizer(device, nr_models):
    import torch.nn as nn
    from dd4ml_b200 import APTS_D
    from dd4ml_b200.workloads import MediumFFNN, apts_d_yaml_config
    torch.manual_seed(42)                       # same initial weights on every rank
    model = MediumFFNN().to(device)
    c = APTS_D.setup_APTS_hparams(apts_d_yaml_config())
    opt = APTS_D(model.parameters(), model=model, criterion=nn.CrossEntropyLoss(), device=device,
                 nr_models=nr_models, glob_opt=c.glob_opt, glob_opt_hparams=c.glob_opt_hparams,
                 loc_opt=c.loc_opt, loc_opt_hparams=c.loc_opt_hparams, glob_pass=c.glob_pass, foc=c.foc,
                 norm_type=c.norm_type, max_loc_iters=c.max_loc_iters, max_glob_iters=c.max_glob_iters,
                 tol=c.tol, **c.apts_params)
    return model, opt


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
        "dd4ml_lssr1_begin": lambda: lib.lssr1_begin(ds.sp, ds.wp, w.data_ptr(), old_w.data_ptr(), gv.data_ptr(),
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
    for name, fn in calls.items():
        if name == "dd4ml_tr_apply":
            lib.tr_propose(ds.sp, ds.wp, gv.data_ptr(), h.S_ptr, h.Y_ptr, n, h.ld, 0, 0, 1, k, st)
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / iters
        kk = int(ds.read().k)
        byt = algorithmic_bytes(name, n, kk)
        out[name] = {"n": n, "k": kk, "ms": ms, "bytes": byt, "achieved": byt / ms / 1e6,
                     "frac": byt / ms / 1e6 / peak_gbs}
    sel = out.get(kernel) or out["dd4ml_lssr1_begin"]
    # DRAM traffic of the same kernel at the same (n, k) from the committed `ncu --set full` capture
    traffic = None
    try:
        key = {"dd4ml_lssr1_begin": "LssrBeginOp", "dd4ml_tr_propose": "ProposeOp", "dd4ml_tr_apply": "ApplyOp",
               "dd4ml_lssr1_eval": "EvalOp", "dd4ml_lssr1_trial": "TrialOp"}[kernel if kernel in out else "dd4ml_lssr1_begin"]
        summ = json.load(open(os.path.join(ROOT, "profiles", "ncu_full_r1_summary.json")))[key]
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tr = 0.0
        for f in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = summ[f].split()
            tr += float(v) * unit[u]
        if n == 1 << 26 and k == 3:
            traffic = tr
    except Exception:  # noqa: BLE001
        traffic = None
    roof = {"bound": "hbm", "achieved": sel["achieved"], "peak": peak_gbs, "unit": "GB/s",
            "frac": sel["frac"], "traffic": traffic, "kernel": kernel if kernel in out else "dd4ml_lssr1_begin",
            "workload": f"fp32 n=2^26 k={sel['k']} (inputs >> L2)", "peak_source": peak_src,
            "bytes_per_launch": sel["bytes"], "ms_per_launch": sel["ms"]}
    return roof, out


def cpu_baseline(steps_budget_s=15.0, n_ranks=1, threads=None):
    """The oracle port of the reference's CPU path, timed on the host cores (reported baseline)."""
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

    fgs = []
    for r in range(n_ranks):
        X, Y = mnist_shape_batch(PER_RANK_BATCH, 1234 + r)
        fgs.append(make_fg(models[r], X, Y))
    ghp = oracle.lssr1_hparams(0.1, 0.01, 1.0, second_order=True, dogleg=True, mem_length=3)
    lhp = oracle.lssr1_loc_hparams(0.1, 0.01, 1.0, n_ranks, second_order=True, dogleg=True, mem_length=3)
    opt = oracle.APTSDOracle(w0, n_ranks, "lssr1_tr", ghp, "lssr1_tr", lhp, **oracle.apts_hparams(0.1, 0.01, 1.0),
                             glob_pass=True, foc=True, norm_type=2, max_loc_iters=3, max_glob_iters=1, tol=1e-6)
    return opt, fgs


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path.  The reference is
    pure Python and cannot travel to the GPU box, so this times the oracle port (pinned against
    the unmodified reference by tests/test_oracle_golden.py) with all host threads."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    n = args.gpus
    threads = os.cpu_count() or 1
    opt, fgs = cpu_baseline(n_ranks=n, threads=threads)
    for _ in range(max(args.warmup, 1)):
        opt.step(fgs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = opt.step(fgs)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{args.steps} outer iterations of the full workload ({n} simulated subdomain(s) in one process)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_loss": float(loss)}))


def workload_config(n):
    return {"workload": "BASELINE configs[1]: MediumFFNN 784-[128]x8-10 (n=217354) on MNIST-shape synthetic data, "
                        "APTS_D glob=loc=LSSR1_TR second-order+dogleg, mem_length 3, max_loc_iters 3, "
                        "max_glob_iters 1, foc, glob_pass",
            "subdomains": n, "per_rank_batch": PER_RANK_BATCH, "parallelism": f"apts_d x{n} (one data subdomain per GPU)",
            "value_definition": "subdomains x outer iterations / s (= outer iterations / s at 1 GPU)",
            "l2_flush": "256 MiB memset before every step, inside the timed region (<1% of a step)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    rank, local, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- dd4ml_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from dd4ml_b200 import _lib
    from dd4ml_b200.workloads import mnist_shape_batch
    lib = _lib.get()
    model, opt = build_optimizer(device, world)
    Xh, Yh = mnist_shape_batch(PER_RANK_BATCH, 1234 + rank, pin=True)      # pinned host batch of this subdomain
    X, Y = Xh.to(device), Yh.to(device)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)

    losses, evals = [], []
    def step_resident(i):
        losses.append(opt.step(X, Y))
        evals.append(float(opt.grad_evals))      # reference counter: global + mean local gradient evaluations

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def sync_count():
        return opt._ctl.syncs + opt.glob_opt._ds.syncs + opt.loc_opt._ds.syncs
    l0, s0 = lib.launches, sync_count()
    del evals[:]
    ms, wall = timed_region(step_resident, args.steps, world, device, flush)
    evals_per_step = sum(evals) / max(len(evals), 1)
    launches = lib.launches - l0
    host_syncs = (sync_count() - s0) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.steps / (ms / 1e3)

    # ---- end-to-end: host batch -> device every step, loss read back every step
    # (the public feed: pinned host batches through HostBatchPrefetcher -- the copy of the next
    # step's batch runs on a side stream under the current step; one H2D copy per step, all of them
    # inside the timed region except the very first)
    import itertools
    from dd4ml_b200.dataloaders import HostBatchPrefetcher
    feed = iter(HostBatchPrefetcher(itertools.repeat((Xh, Yh)), device))
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    def step_e2e(i):
        Xd, Yd = next(feed)
        loss = opt.step(Xd, Yd)
        host_loss.copy_(loss.detach().float(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed_region(step_e2e, args.steps, world, device, flush)
    e2e = {"value": world * args.steps / (ms_e2e / 1e3), "unit": UNIT,
           "h2d_bytes_per_step": Xh.numel() * Xh.element_size() + Yh.numel() * Yh.element_size(),
           "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps}

    # ---- per-kernel device times inside the step (CUDA events around every C-ABI launch)
    lib.prof = {}
    nprof = min(args.steps, 10)
    for i in range(nprof):
        flush.zero_()
        step_resident(i)
    torch.cuda.synchronize(device)
    prof, lib.prof = lib.prof, None
    n = sum(p.numel() for p in model.parameters())
    kmem = 3
    native = {}
    for name, evs in prof.items():
        ts = [a.elapsed_time(b) for a, b in evs]
        byt = algorithmic_bytes(name, n, kmem)
        native[name] = {"launches_per_step": len(ts) / nprof, "avg_us": 1e3 * sum(ts) / len(ts),
                        "min_us": 1e3 * min(ts), "median_us": 1e3 * statistics.median(ts), "max_us": 1e3 * max(ts),
                        "total_ms_per_step": sum(ts) / nprof,
                        "achieved_gbs": None if not byt else byt / (sum(ts) / len(ts)) / 1e6}
    dominant = max(native, key=lambda k: native[k]["total_ms_per_step"]) if native else "dd4ml_lssr1_begin"

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    roofline, sweep = (None, None)
    if not args.no_sweep:
        roofline, sweep = large_n_roofline(device, dominant, peak, peak_src)
    own_ms = sum(v["total_ms_per_step"] for v in native.values())
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        copt, cfgs = cpu_baseline(threads=os.cpu_count() or 1)
        copt.step(cfgs)
        t0, it = time.perf_counter(), 0
        while time.perf_counter() - t0 < 15.0 and it < 200:
            copt.step(cfgs); it += 1
        dt = time.perf_counter() - t0
        cpu = {"value": it / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{it} outer iterations of the same workload after 1 warm-up ({dt:.1f} s)"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "outer_iters_per_sec": args.steps / (ms / 1e3), "wall_ms_per_step": wall * 1e3 / args.steps,
        "e2e": e2e, "gpu_launches": launches, "host_syncs_per_step": host_syncs,
        # work per outer iteration is trajectory dependent (line-search evaluations): more subdomains
        # take more evaluations per iteration, which the per-N ms_per_step reflects
        "grad_evals_per_step": evals_per_step,
        "clocks": clocks, "roofline": roofline,
        "roofline_native": {"n": n, "k": kmem, "dominant": dominant, "own_kernel_ms_per_step": own_ms,
                            "own_kernel_share_of_step": own_ms / (ms / args.steps), "kernels": native,
                            "note": "native vectors are 0.87 MB (L2 resident): launch-latency-bound, see DESIGN.md"},
        "roofline_sweep": sweep, "cpu_baseline": cpu,
        "final_loss": float(losses[-1]), "first_loss": float(losses[0]),
    }
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
