"""Diagnostics (GPU): k-space finalize cycles of lssr1_begin back to back vs. after unrelated kernels
(instruction / constant cache cold), at the bench-native and at a large vector length."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dd4ml_b200 import _lib
from dd4ml_b200.optimizers.lsr1 import LSR1
dev = torch.device("cuda", 0)
lib = _lib.get()
for n in (217354, 1 << 24):
    k = 3
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
    A = torch.randn(4096, 4096, device=dev)
    def begin():
        lib.lssr1_begin(ds.sp, ds.wp, w.data_ptr(), old_w.data_ptr(), old_w.data_ptr(), gv.data_ptr(), prev_g.data_ptr(), prev_g.data_ptr(), 0,
                        h.S_ptr, h.Y_ptr, loss.data_ptr(), 0, n, h.ld, 0, 0, 0, k, -1.0, 0.0, 0.0, _lib.stream())
        rep = ds.read()
        return int(rep.cyc_solve), int(rep.cyc_eigh), int(rep.cyc_newton)
    print(f"n={n}: back to back  ", [begin() for _ in range(4)])
    out = []
    for _ in range(4):
        for _ in range(3):
            (A @ A).relu_().sum()
            torch.nn.functional.softmax(A, dim=1)
        out.append(begin())
    print(f"n={n}: after other kernels", out)
