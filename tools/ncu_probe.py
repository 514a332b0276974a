"""GPU: a few launches of one heavy kernel at n = 2^26 for an `ncu --set full` capture (SWEEP_K pairs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import large_n_roofline
k = int(os.environ.get("SWEEP_K", 10))
roof, out = large_n_roofline(torch.device("cuda", 0), "dd4ml_tr_propose", 6459.3, "measured", n=1 << 26, k=k, iters=2)
for name, v in out.items():
    print(f"{name:22s} k={v['k']} {v['ms']:.4f} ms frac {v['frac']:.3f}")
