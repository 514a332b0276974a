"""Diagnostics (GPU): cycle counters of the k-space solve inside tr_propose for k = 1..10
(well-conditioned synthetic memory, n = 50 000), to see where the single-thread time goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dd4ml_b200 import _lib
from dd4ml_b200.optimizers.lsr1 import LSR1
dev = torch.device("cuda", 0)
lib = _lib.get()
n = 50000
torch.manual_seed(0)
for k in (1, 2, 3, 4, 5, 6, 8, 10):
    h = LSR1(gamma=1.0, memory_length=k, tol=1e-6, device=dev)
    d = torch.linspace(0.5, 2.0, n, device=dev)
    for _ in range(k):
        s = torch.randn(n, device=dev) * 0.01
        h.update_memory(s, d * s)
    ds = h.ds
    g = torch.randn(n, device=dev)
    for delta, dog in ((0.05, 0.0), (1.0, 1.0), (100.0, 0.0)):
        ds.set(delta=delta, min_delta=1e-3, max_delta=1e3, tol=1e-6, second_order=1.0, dogleg=dog)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.tr_propose(ds.sp, ds.wp, g.data_ptr(), h.S_ptr, h.Y_ptr, n, h.ld, 0, 0, 1, k, _lib.stream())
        e0.record()
        lib.tr_propose(ds.sp, ds.wp, g.data_ptr(), h.S_ptr, h.Y_ptr, n, h.ld, 0, 0, 1, k, _lib.stream())
        e1.record()
        r = ds.read()
        print(f"k={k:2d} delta={delta:6.2f} dog={int(dog)} case={int(r.case_id)} kernel {e0.elapsed_time(e1) * 1e3:6.1f} us  "
              f"cyc solve {int(r.cyc_solve):7d} eigh {int(r.cyc_eigh):7d} newton {int(r.cyc_newton):7d} "
              f"newton_iters {int(r.newton_iters):3d} sweeps {int(r.jacobi_sweeps)}")
