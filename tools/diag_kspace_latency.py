"""Diagnostics (GPU): latency of the entry points that end in the one-CTA k-space kernel
(lssr1_begin, tr_propose) at the bench-native vector length, for k = 3 / 5 / 10 stored pairs:
back to back (instruction cache warm) and right after unrelated kernels (the closure's forward /
backward: instruction cache cold), CUDA events around the single call, plus the solver's own
clock64 counters.  Prints a markdown table (committed under profiles/)."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dd4ml_b200 import _lib
from dd4ml_b200.optimizers.lsr1 import LSR1

dev = torch.device("cuda", 0)
lib = _lib.get()
A = torch.randn(4096, 4096, device=dev)
B = torch.randn(2048, 784, device=dev)
W1 = torch.randn(784, 128, device=dev)


def evict():
    for _ in range(3):
        (A @ A).relu_().sum()
        torch.nn.functional.log_softmax(torch.relu(B @ W1), dim=1).sum()


def timed(fn, cold, reps=12):
    ts, cyc = [], []
    for _ in range(reps):
        if cold:
            evict()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
        cyc.append(rep_of())
    return ts, cyc


rows = []
for n in (217354, 2719104):
    for k in (3, 5, 10):
        h = LSR1(gamma=1e-3, memory_length=k, tol=1e-6, device=dev)
        g = torch.Generator(device=dev).manual_seed(1)
        for _ in range(k):
            s = torch.randn(n, device=dev, generator=g) * 0.01
            h.update_memory(s, s + 0.001 * torch.randn(n, device=dev, generator=g))
        ds = h.ds
        ds.set(delta=0.1, min_delta=0.01, max_delta=1.0, tol=1e-6, second_order=1.0, dogleg=1.0)
        r = lambda: torch.randn(n, device=dev, generator=g)  # noqa: E731
        w, old_w, gv, prev_g = r(), r(), r(), r()
        loss = torch.tensor(1.0, device=dev)

        def rep_of():
            rp = ds.read()
            return int(rp.cyc_solve), int(rp.cyc_eigh), int(rp.cyc_newton), int(rp.newton_iters), int(rp.k)

        def begin():
            lib.lssr1_begin(ds.sp, ds.wp, w.data_ptr(), old_w.data_ptr(), old_w.data_ptr(), gv.data_ptr(), prev_g.data_ptr(), prev_g.data_ptr(), 0,
                            h.S_ptr, h.Y_ptr, loss.data_ptr(), 0, n, h.ld, 0, 0, 0, k, -1.0, 0.0, 0.0, _lib.stream())

        def propose():
            lib.tr_propose(ds.sp, ds.wp, gv.data_ptr(), h.S_ptr, h.Y_ptr, n, h.ld, 0, 0, 0, k, _lib.stream())

        def empty():
            pass
        for name, fn in (("lssr1_begin", begin), ("tr_propose", propose)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            tw, cw = timed(fn, cold=False)
            tc, cc = timed(fn, cold=True)
            t0, _ = timed(empty, cold=False)
            rows.append((n, k, name, statistics.median(tw), min(tw), statistics.median(tc), min(tc), max(tc),
                         statistics.median(t0), statistics.median(c[0] for c in cw), statistics.median(c[0] for c in cc),
                         statistics.median(c[1] for c in cc), statistics.median(c[2] for c in cc), cc[0][3], cc[0][4]))

print("| n | k | entry | warm us (median / min) | cold us (median / min / max) | empty event pair us | "
      "cyc_solve warm | cyc_solve cold | cyc_eigh cold | cyc_newton cold | newton iters | live pairs |")
print("|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.1f} / {r[4]:.1f} | {r[5]:.1f} / {r[6]:.1f} / {r[7]:.1f} | {r[8]:.1f} | "
          f"{r[9]:.0f} | {r[10]:.0f} | {r[11]:.0f} | {r[12]:.0f} | {r[13]} | {r[14]} |")
