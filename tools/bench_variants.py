"""GPU: the bench workload (MediumFFNN n = 217 354, 5000 samples, one subdomain) with the other
optimizer pairings of the reference's config space, to show the sync-free TR/TR path next to the
LSSR1_TR/LSSR1_TR headline: outer iterations/s, host reads and kernel launches per iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from bench import PER_RANK_BATCH, timed_region
from dd4ml_b200 import APTS_D, _lib
from dd4ml_b200.workloads import MediumFFNN, apts_d_yaml_config, mnist_shape_batch

dev = torch.device("cuda", 0)
X, Y = mnist_shape_batch(PER_RANK_BATCH, 1234, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.get()
print("| glob / loc | order | mem | outer it/s | ms / it | host reads / it | dd4ml_b200 launches / it | grad evals / it |")
print("|---|---|---:|---:|---:|---:|---:|---:|")
for glob, loc, so in (("lssr1_tr", "lssr1_tr", True), ("lssr1_tr", "lssr1_tr", False), ("tr", "tr", True),
                      ("tr", "tr", False), ("lssr1_tr", "tr", True)):
    torch.manual_seed(42)
    model = MediumFFNN().to(dev)
    c = apts_d_yaml_config()
    c.glob_opt, c.loc_opt = glob, loc
    c.glob_second_order = c.loc_second_order = so
    c.glob_dogleg = c.loc_dogleg = so
    c = APTS_D.setup_APTS_hparams(c)
    opt = APTS_D(model.parameters(), model=model, criterion=nn.CrossEntropyLoss(), device=dev, nr_models=1,
                 glob_opt=c.glob_opt, glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt,
                 loc_opt_hparams=c.loc_opt_hparams, glob_pass=c.glob_pass, foc=c.foc, norm_type=c.norm_type,
                 max_loc_iters=c.max_loc_iters, max_glob_iters=c.max_glob_iters, tol=c.tol, **c.apts_params)
    ev = []
    def step(i):
        opt.step(X, Y); ev.append(float(opt.grad_evals))
    for i in range(5):
        step(i)
    syncs = lambda: opt._ctl.syncs + opt.glob_opt._ds.syncs + opt.loc_opt._ds.syncs  # noqa: E731
    s0, l0 = syncs(), lib.launches
    del ev[:]
    steps = 30
    ms, _ = timed_region(step, steps, 1, dev, flush)
    mem = c.glob_opt_hparams.get("mem_length", "-") if so else "-"
    print(f"| {glob} / {loc} | {'2nd + dogleg' if so else '1st'} | {mem} | {steps / ms * 1e3:.1f} | {ms / steps:.2f} | "
          f"{(syncs() - s0) / steps:.1f} | {(lib.launches - l0) / steps:.1f} | {sum(ev) / len(ev):.1f} |", flush=True)
