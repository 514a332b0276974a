"""GPU: the heavy entry points at n = 2^26 (+ pad), k = CHECK_K pairs, checked against torch on the same data:
state flags after every call and the direction written by tr_apply (p = cg g + sum_j cS_j S_j + cY_j Y_j)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dd4ml_b200 import _lib
from dd4ml_b200.optimizers.lsr1 import LSR1

dev = torch.device("cuda", 0)
lib = _lib.get()
k = int(os.environ.get("CHECK_K", 10))
for n in (int(os.environ.get("CHECK_N", 1 << 26)), (1 << 26) + 8192):
    h = LSR1(gamma=1e-3, memory_length=k, tol=1e-6, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    for _ in range(k):
        s = torch.randn(n, device=dev, generator=g) * 0.01
        y = s + 0.1 * torch.randn(n, device=dev, generator=g) * 0.01
        h.update_memory(s, y)
    ds = h.ds
    ds.set(delta=0.1, min_delta=0.01, max_delta=1.0, tol=1e-6, second_order=1.0, dogleg=1.0)
    gv = torch.randn(n, device=dev, generator=g)
    p = torch.empty(n, device=dev)
    st = _lib.stream()
    lib.tr_propose(ds.sp, ds.wp, gv.data_ptr(), h.S_ptr, h.Y_ptr, n, h.ld, 0, 0, 1, k, st)
    rep = ds.read(full=True)
    print(f"n={n} k={int(rep.k)} err={rep.err} valid={rep.solve_valid} case={rep.case_id} sweeps={rep.jacobi_sweeps} "
          f"newton={rep.newton_iters} cyc_solve={rep.cyc_solve:.0f} cyc_eigh={rep.cyc_eigh:.0f} cyc_newton={rep.cyc_newton:.0f} ld={h.ld}")
    lib.tr_apply(ds.sp, ds.wp, gv.data_ptr(), h.S_ptr, h.Y_ptr, p.data_ptr(), None, n, h.ld, 0, 0, 0, k, st)
    ref = rep.cg * gv.double()
    order = [int(o) for o in rep.order[:int(rep.k)]]
    Sbuf, Ybuf = h._Sbuf, h._Ybuf
    for sl in order:
        ref += rep.cS[sl] * Sbuf[sl, :n].double() + rep.cY[sl] * Ybuf[sl, :n].double()
    err = float((p.double() - ref).norm() / ref.norm())
    print(f"   tr_apply direction vs torch: rel err {err:.2e}  |p|={float(p.norm()):.4e} pnorm(state)={rep.pnorm:.4e}")
    del h, p, gv, ref
    torch.cuda.empty_cache()
