"""GPU: SURVEY §8(d) size sweep of the fused kernels: n in {native, 2^20 .. 2^28} x k in {3, 5, 10}.
Prints one markdown table row per (n, k): fraction of the measured copy peak per kernel
(CUDA events, 10 launches each, no profiler).  Working sets below ~126 MB stay in L2 between
launches (marked L2): those rows measure launch latency / L2 bandwidth, not HBM."""
import gc
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import large_n_roofline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
names = ["dd4ml_lssr1_begin", "dd4ml_tr_propose", "dd4ml_tr_apply", "dd4ml_tr_decide", "dd4ml_lssr1_eval",
         "dd4ml_lssr1_trial"]
print("Cells: entry point us (streaming kernel + k-space kernel) / streaming kernel alone us, fraction of the "
      "measured copy peak of the streaming kernel.\n")
print("| n | k | working set | " + " | ".join(n.replace("dd4ml_", "") for n in names) + " |")
print("|---:|---:|---|" + "---:|" * len(names))
for n in [217354, 1 << 20, 1 << 22, 1 << 24, 1 << 26, 1 << 28]:
    for k in (3, 5, 10):
        _, out = large_n_roofline(torch.device("cuda", 0), "dd4ml_lssr1_begin", peak, "measured", n=n, k=k, iters=10)
        ws = (2 * k + 8) * 4 * n
        cells = []
        for nm in names:
            v = out.get(nm)
            cells.append("—" if v is None else
                         f"{v['ms_entry_with_kspace_kernel'] * 1e3:.1f} / {v['ms'] * 1e3:.1f}, {v['frac']:.2f}")
        tag = f"{ws / 1e6:.0f} MB" + (" (L2)" if ws < 120e6 else "")
        ln = f"2^{n.bit_length() - 1}" if n & (n - 1) == 0 else str(n)
        print(f"| {ln} | {k} | {tag} | " + " | ".join(cells) + " |", flush=True)
        gc.collect()
        torch.cuda.empty_cache()
