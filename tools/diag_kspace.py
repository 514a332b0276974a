"""Diagnostics (GPU): k-space cycle counters of the LSSR1 local/global steps of the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_optimizer, PER_RANK_BATCH
from dd4ml_b200.workloads import mnist_shape_batch
dev = torch.device("cuda", 0)
model, opt = build_optimizer(dev, 1)
X, Y = mnist_shape_batch(PER_RANK_BATCH, 1234, device=dev)
for i in range(12):
    loss = opt.step(X, Y)
    lr, gr = opt.loc_opt.last_report, opt.glob_opt.last_report
    f = lambda r: {k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items() if k in
                   ("cyc_solve", "cyc_eigh", "cyc_newton", "newton_iters", "jacobi_sweeps", "k", "case", "evals", "alpha")}
    print(i, round(float(loss), 5), "loc", f(lr), "glob", f(gr))
