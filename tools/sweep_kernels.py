"""GPU: launch the heavy fused kernels on vectors >> L2 (for ncu captures and quick A/B timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import large_n_roofline
peak = 6459.3
n = int(os.environ.get("SWEEP_N", 1 << 26))
k = int(os.environ.get("SWEEP_K", 3))
roof, out = large_n_roofline(torch.device("cuda", 0), "dd4ml_lssr1_begin", peak, "measured", n=n, k=k, iters=int(os.environ.get("SWEEP_ITERS", 10)))
for name, v in out.items():
    print(f"{name:22s} n={v['n']} k={v['k']} {v['ms']:.4f} ms  {v['achieved']:.0f} GB/s  frac {v['frac']:.3f}   "
          f"(entry point incl. k-space kernel: {v['ms_entry_with_kspace_kernel']:.4f} ms, frac "
          f"{v['frac_entry_with_kspace_kernel']:.3f})")
