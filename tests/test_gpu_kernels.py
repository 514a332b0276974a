"""GPU parity tests of the fused kernels, called through the product API (which goes
through the C ABI), against the CPU oracle and the reference goldens."""
import math

import numpy as np
import pytest
import torch

from _helpers import oracle_memory, sigma_spread, load_golden, rel_err

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import dd4ml_b200
    from dd4ml_b200 import _lib
    from dd4ml_b200._device import DeviceState
    from dd4ml_b200.optimizers.lsr1 import LSR1
    from dd4ml_b200.utility.optimizer_utils import solve_tr_first_order, solve_tr_second_order

import oracle
from oracle.lsr1 import LSR1Memory
from oracle.subproblem import solve_first_order, solve_second_order

DEV = "cuda"


@pytest.mark.parametrize("n", [1, 3, 1000, 1023, 217354, (1 << 22) + 5])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("norm", [2, math.inf])
def test_first_order_solve(n, dtype, norm):
    torch.manual_seed(n)
    g = torch.randn(n, dtype=dtype)
    gn = torch.norm(g, p=norm)
    p, pred = solve_tr_first_order(g.to(DEV), gn, 0.37, 1e-6)
    # exact arithmetic (fp64 oracle on the same fp32 data): the kernel reduces in fp64
    g64 = g.double()
    p64, pred64 = solve_first_order(g64, torch.norm(g64, p=norm), 0.37, 1e-6)
    assert rel_err(p.cpu(), p64) < 2e-7 if dtype == torch.float32 else rel_err(p.cpu(), p64) < 1e-14
    assert float(pred) == pytest.approx(float(pred64), rel=2e-7)
    # the reference's own dtype: torch's fp32 CPU norm carries ~1e-6 (n = 2e5) ... 1e-4 (n = 4e6)
    # accumulation error, so the 1e-5 north-star tolerance is asserted at the BASELINE sizes only
    p_ref, pred_ref = solve_first_order(g, gn, 0.37, 1e-6)
    if n <= 217354:
        assert rel_err(p.cpu(), p_ref) < 1e-5
        assert float(pred) == pytest.approx(float(pred_ref), rel=1e-5)


def test_first_order_unaligned_pointer():
    torch.manual_seed(0)
    base = torch.randn(4099, device=DEV)
    g = base[1:]                       # 4-byte aligned only -> scalar path
    p, pred = solve_tr_first_order(g, torch.norm(g), 0.5, 1e-6)
    p_ref, pred_ref = solve_first_order(g.cpu(), torch.norm(g.cpu()), 0.5, 1e-6)
    assert rel_err(p.cpu(), p_ref) < 1e-6


def _replay_updates(gold, m, hess_gpu, hess_cpu, idx0):
    idx = idx0
    for _ in range(m + 2):
        s, y = torch.from_numpy(gold[f"upd{idx}_s"]), torch.from_numpy(gold[f"upd{idx}_y"])
        hess_gpu.update_memory(s.to(DEV, hess_gpu.dtype), y.to(DEV, hess_gpu.dtype))
        if hess_cpu is not None:
            hess_cpu.update(s, y)
            hess_cpu.precompute()
        yield idx
        idx += 1


def test_lsr1_update_chain_gamma_and_B_match_reference():
    gold = load_golden("kernel_level")
    idx0 = 0
    for m in (1, 3, 5):
        h = LSR1(gamma=1.0, memory_length=m, tol=1e-6, device=DEV)
        hc = LSR1Memory(gamma=1.0, memory_length=m, tol=1e-6)
        for idx in _replay_updates(gold, m, h, hc, idx0):
            assert float(h.gamma) == pytest.approx(float(gold[f"upd{idx}_gamma"]), rel=1e-5)
            assert len(h._S) == len(hc)
            v = torch.from_numpy(gold[f"upd{idx}_v"])
            Bv = h.B(v.to(DEV)).cpu()
            assert rel_err(Bv, gold[f"upd{idx}_Bv"]) < 1e-5
            assert rel_err(h.Minv.cpu(), gold[f"upd{idx}_Minv"]) < 1e-4
        # ring order == reference column order after wrap-around
        assert rel_err(h.S.cpu(), torch.stack(hc.S, 1)) == 0.0
        assert rel_err(h.Y.cpu(), torch.stack(hc.Y, 1)) == 0.0
        idx0 += m + 2


@pytest.mark.parametrize("m", [1, 3, 5])
@pytest.mark.parametrize("dog", [False, True])
@pytest.mark.parametrize("delta", [0.01, 0.3, 5.0, 100.0])
def test_second_order_solve(m, dog, delta):
    gold = load_golden("kernel_level")
    idx0 = {1: 0, 3: 3, 5: 8}[m]
    h = LSR1(gamma=1.0, memory_length=m, tol=1e-6, device=DEV)
    hc = LSR1Memory(gamma=1.0, memory_length=m, tol=1e-6)
    for _ in _replay_updates(gold, m, h, hc, idx0):
        pass
    key = f"so_m{m}_dog{int(dog)}_d{delta}"
    g = torch.from_numpy(gold[key + "_g"])
    obs = dd4ml_b200.OBS()
    p, pred = solve_tr_second_order(g.to(DEV), torch.norm(g), delta, h, obs, 1e-6, dogleg=dog)
    info = {}
    p_or, pred_or = solve_second_order(g, torch.norm(g), delta, hc, 1e-6, dogleg=dog, eig_sign="canonical", info=info)
    assert {"A": 1, "C": 3}[info["case"]] == obs.last["case"]
    # Yardstick: the float64 evaluation of the reference algorithm on the same pairs.  The reference's own
    # float32 result (the golden) sits e_ref away from it: <= 4e-7 on 19 of these 24 solves and 2e-6 ... 1.7e-4
    # on the delta = 100 ones (cond(Psi^T Psi) * eps of its float32 LAPACK pipeline).  Stated bound: the CUDA
    # path is within 1e-5 of the REFERENCE GOLDEN wherever the reference itself is that accurate, and within
    # the reference's own float32 error elsewhere.  None of the 24 solves is eigenvector-sign sensitive
    # (tests/test_oracle_golden.py::test_sign_sensitivity_census_of_the_reference_goldens).
    hc64 = oracle_memory(torch.stack(hc.S, 1), torch.stack(hc.Y, 1), float(hc.gamma), torch.float64)
    p64, pred64 = solve_second_order(g.double(), torch.norm(g.double()), delta, hc64, 1e-6, dogleg=dog, eig_sign="canonical")
    e_ref = rel_err(gold[key + "_p"], p64)
    assert sigma_spread(info, delta, float(hc.gamma)) <= 1e-6
    assert rel_err(p.cpu(), gold[key + "_p"]) < 1e-5 + 1.5 * e_ref, (e_ref, rel_err(p.cpu(), gold[key + "_p"]))
    assert rel_err(p.cpu(), p_or) < 1e-5 + 1.5 * e_ref
    pe_ref = abs(float(gold[key + "_pred"]) - float(pred64)) / max(abs(float(pred64)), 1e-12)
    # (pred also carries sigma*, which the reference iterates in float32 and the yardstick in float64: factor 3)
    assert float(pred) == pytest.approx(float(gold[key + "_pred"]), rel=1e-5 + 3 * pe_ref, abs=1e-9)


@pytest.mark.parametrize("m", [8, 10])
@pytest.mark.parametrize("dog", [False, True])
@pytest.mark.parametrize("delta", [0.01, 0.3, 5.0, 100.0])
def test_second_order_solve_at_the_reference_default_memory_length(m, dog, delta):
    """8 and 10 stored pairs (10 = the reference's default memory_length) through the KMAX = 8 / 10 kernels and
    the four-warp k-space kernel, against the UNMODIFIED reference (tests/golden/kernel_level_m8_m10.npz): the
    reference's update chain is replayed pair by pair (acceptance, gamma, ring order), then the 12 solves that do
    not depend on LAPACK's eigenvector signs are compared with the reference's float32 step at 1e-5 + 1.5 e_ref;
    the 4 sign-sensitive delta = 100 solves are run in float64 and compared with the float64 oracle under the
    documented sign convention (same rule as tests/test_kspace_host.py)."""
    gold = load_golden("kernel_level_m8_m10")
    idx0 = {8: 0, 10: 10}[m]
    key = f"so_m{m}_dog{int(dog)}_d{delta}"
    g = torch.from_numpy(gold[key + "_g"])
    S, Y, gamma = torch.from_numpy(gold[key + "_S"]), torch.from_numpy(gold[key + "_Y"]), float(gold[key + "_gamma"])
    hc64 = oracle_memory(S, Y, gamma, torch.float64)
    p64, pred64 = solve_second_order(g.double(), torch.norm(g.double()), delta, hc64, 1e-6, dogleg=dog, eig_sign="canonical")
    info = {}
    solve_second_order(g, torch.norm(g), delta, oracle_memory(S, Y, gamma, torch.float32), 1e-6, dogleg=dog,
                       eig_sign="canonical", info=info)
    sensitive = sigma_spread(info, delta, gamma) > 1e-6
    assert sensitive == (delta == 100.0)
    dtype = torch.float64 if sensitive else torch.float32
    h = LSR1(gamma=1.0, memory_length=m, tol=1e-6, device=DEV, dtype=dtype)
    for idx in _replay_updates(gold, m, h, None, idx0):
        assert float(h.gamma) == pytest.approx(float(gold[f"upd{idx}_gamma"]), rel=1e-5)
    assert len(h._S) == m
    assert rel_err(h.S.cpu().float(), S) == 0.0 and rel_err(h.Y.cpu().float(), Y) == 0.0      # ring order == reference columns
    obs = dd4ml_b200.OBS()
    gd = g.to(DEV, dtype)
    p, pred = solve_tr_second_order(gd, torch.norm(gd), delta, h, obs, 1e-6, dogleg=dog)
    if not sensitive:
        e_ref = rel_err(gold[key + "_p"], p64)
        assert rel_err(p.cpu(), gold[key + "_p"]) < 1e-5 + 1.5 * e_ref, (e_ref, rel_err(p.cpu(), gold[key + "_p"]))
        assert float(pred) == pytest.approx(float(gold[key + "_pred"]), rel=2e-5, abs=1e-9)
    else:
        # the float64 run derives gamma from the float32 pairs in float64 (1e-7 from the reference's float32
        # gamma): the yardstick uses the same gamma
        hcg = oracle_memory(S, Y, float(h.gamma), torch.float64)
        pg, predg = solve_second_order(g.double(), torch.norm(g.double()), delta, hcg, 1e-6, dogleg=dog, eig_sign="canonical")
        assert rel_err(p.cpu(), pg) < 1e-6
        assert float(pred) == pytest.approx(float(predg), rel=1e-6, abs=1e-12)


GOLDEN_TRAJ = ("tr_so", "tr_so_dogleg", "lssr1_so_dog0", "lssr1_so_dog1", "lssr1_paper")


@pytest.mark.parametrize("name", GOLDEN_TRAJ)
def test_solves_recorded_inside_reference_trajectories(name):
    """The 8 sub-problem solves the unmodified reference performed first in each second-order trajectory
    golden (inputs g, S, Y, gamma, delta and output p, pred recorded by gen_golden.SolveRecorder), re-solved
    on the GPU from the recorded inputs.  Per solve: (i) sign-insensitive (39 of 40): compared with the
    REFERENCE's p at 1e-5 + 1.5 e_ref, e_ref = the reference's own distance from the float64 evaluation of its
    algorithm (<= 1e-5 on 31 solves, up to 8e-3 on the ill-conditioned ones); (ii) the one sign-sensitive
    solve (lssr1_so_dog1 #6): compared with the oracle under this implementation's documented sign
    convention, same bound."""
    from _helpers import golden_solves
    obs = dd4ml_b200.OBS()
    n_sens = 0
    for i, r in enumerate(golden_solves(name)):
        g = torch.from_numpy(r["g"])
        delta, dog, tol, gamma = float(r["delta"]), bool(int(r["dogleg"])), float(r["tol"]), float(r["gamma"])
        k = r["S"].shape[1]
        h = LSR1(gamma=gamma, memory_length=max(k, 1), tol=tol, device=DEV)
        h.set_memory(torch.from_numpy(r["S"]), torch.from_numpy(r["Y"]), gamma)
        p, pred = solve_tr_second_order(g.to(DEV), torch.norm(g), delta, h, obs, tol, dogleg=dog)
        info = {}
        p_can, pred_can = solve_second_order(g, torch.norm(g), delta, oracle_memory(r["S"], r["Y"], gamma), tol,
                                             dogleg=dog, eig_sign="canonical", info=info)
        p64, _ = solve_second_order(g.double(), torch.norm(g.double()), delta,
                                    oracle_memory(r["S"], r["Y"], gamma, torch.float64), tol, dogleg=dog,
                                    eig_sign="canonical")
        sens = sigma_spread(info, delta, gamma) > 1e-6
        n_sens += sens
        target = p_can if sens else torch.from_numpy(r["p"])
        e_ref = rel_err(target, p64)
        err = rel_err(p.cpu(), target)
        print(f"{name} solve {i}: k={k} case={info['case']} sensitive={sens} e_ref={e_ref:.1e} gpu-vs-{'oracle' if sens else 'reference'}={err:.1e}")
        assert err < 1e-5 + 1.5 * e_ref, (name, i, err, e_ref)
        assert {"A": 1, "C": 3}[info["case"]] == obs.last["case"]
    assert n_sens == (1 if name == "lssr1_so_dog1" else 0)


def test_second_order_well_conditioned_within_1e5():
    """North-star tolerance (1e-5 relative, fp32) on a well-conditioned memory."""
    torch.manual_seed(3)
    n, m = 50000, 3
    h = LSR1(gamma=1.0, memory_length=m, tol=1e-6, device=DEV)
    hc = LSR1Memory(gamma=1.0, memory_length=m, tol=1e-6)
    d = torch.linspace(0.5, 2.0, n)
    for _ in range(m):
        s = torch.randn(n) * 0.01
        y = d * s
        h.update_memory(s.to(DEV), y.to(DEV))
        assert hc.update(s, y)
        hc.precompute()
    assert len(h._S) == m
    for delta in (0.05, 1.0):
        g = torch.randn(n)
        p, pred = solve_tr_second_order(g.to(DEV), torch.norm(g), delta, h, None, 1e-6, dogleg=False)
        p_or, pred_or = solve_second_order(g, torch.norm(g), delta, hc, 1e-6, eig_sign="canonical")
        assert rel_err(p.cpu(), p_or) < 1e-5
        assert float(pred) == pytest.approx(float(pred_or), rel=1e-4)


@pytest.mark.parametrize("m", [6, 8, 10])
def test_memory_lengths_up_to_the_reference_default(m):
    """LSR1 / TR default to memory_length = 10 in the reference (lsr1.py:24, tr.py:40): the KMAX = 10
    kernel variants, with eviction (m + 2 pairs pushed), against the oracle."""
    torch.manual_seed(5)
    n = 40000 + m
    h = LSR1(gamma=1.0, memory_length=m, tol=1e-6, device=DEV)
    hc = LSR1Memory(gamma=1.0, memory_length=m, tol=1e-6)
    d = torch.linspace(0.5, 2.0, n)
    for i in range(m + 2):
        s = torch.randn(n) * 0.01
        y = d * s
        h.update_memory(s.to(DEV), y.to(DEV))
        assert hc.update(s, y)
        hc.precompute()
    assert len(h._S) == m == len(hc)
    assert rel_err(h.S.cpu(), torch.stack(hc.S, 1)) == 0.0
    v = torch.randn(n)
    assert rel_err(h.B(v.to(DEV)).cpu(), hc.B(v)) < 1e-5
    for delta, dog in ((0.05, False), (1.0, True)):
        g = torch.randn(n)
        p, pred = solve_tr_second_order(g.to(DEV), torch.norm(g), delta, h, None, 1e-6, dogleg=dog)
        p_or, pred_or = solve_second_order(g, torch.norm(g), delta, hc, 1e-6, dogleg=dog, eig_sign="canonical")
        assert rel_err(p.cpu(), p_or) < 2e-5
        assert float(pred) == pytest.approx(float(pred_or), rel=1e-4)
    with pytest.raises(dd4ml_b200.DD4MLError):
        LSR1(gamma=1.0, memory_length=11, tol=1e-6, device=DEV)


def test_cholesky_failure_raises_like_torch():
    n = 256
    h = LSR1(gamma=1.0, memory_length=3, tol=1e-6, device=DEV)
    torch.manual_seed(0)
    s = torch.randn(n, device=DEV)
    h.update_memory(s, 2.0 * s + 0.1 * torch.randn(n, device=DEV))
    # duplicate the stored pair by hand -> Psi^T Psi singular
    rep = h.ds.read(full=True)
    ms = h.ds.lay.maxs
    h._Sbuf[1].copy_(h._Sbuf[0]); h._Ybuf[1].copy_(h._Ybuf[0])
    SS = np.array(rep.SS).reshape(ms, ms); SY = np.array(rep.SY).reshape(ms, ms); YY = np.array(rep.YY).reshape(ms, ms)
    for A in (SS, SY, YY):
        A[0, 1] = A[1, 0] = A[1, 1] = A[0, 0]
    YY[1, 1] = -abs(YY[0, 0])      # make Psi^T Psi indefinite beyond rounding
    h.ds.set_array("SS", SS.ravel()); h.ds.set_array("SY", SY.ravel()); h.ds.set_array("YY", YY.ravel())
    h.ds.set_array("order", [0, 1]); h.ds.set(k=2, spare=2)
    with pytest.raises(torch.linalg.LinAlgError):
        solve_tr_second_order(torch.randn(n, device=DEV), 1.0, 0.1, h, None, 1e-6)


def test_apts_vector_kernels_match_torch():
    """residual / foc / pack / trial / control against plain torch on the same inputs."""
    lib = _lib.get()
    torch.manual_seed(1)
    for n in (5, 1000, 100003):
        for dtype in (torch.float32, torch.float64):
            wt = _lib.dt(dtype)
            st = _lib.stream()
            r = lambda: torch.randn(n, dtype=dtype, device=DEV)  # noqa: E731
            Gg, Gl = r(), r()
            g0, gl0, resid = torch.empty_like(Gg), torch.empty_like(Gg), torch.empty_like(Gg)
            lib.apts_residual(Gg.data_ptr(), Gl.data_ptr(), g0.data_ptr(), gl0.data_ptr(), resid.data_ptr(), n, wt, st)
            assert torch.equal(g0, Gg) and torch.equal(gl0, Gl) and torch.equal(resid, Gg - Gl)
            ds = DeviceState(DEV)
            wl, wg = r(), r()
            loss = torch.tensor(1.25, dtype=torch.float32, device=DEV)
            out = torch.empty_like(loss)
            lib.apts_foc(ds.sp, ds.wp, wl.data_ptr(), wg.data_ptr(), resid.data_ptr(), loss.data_ptr(), out.data_ptr(),
                         0, 1e-6, n, wt, wt, st)
            ref = loss.double() + (resid.double() @ (wl - wg).double())
            assert float(out) == pytest.approx(float(ref), rel=2e-6)
            lib.apts_foc(ds.sp, ds.wp, wl.data_ptr(), wl.data_ptr(), resid.data_ptr(), loss.data_ptr(), out.data_ptr(),
                         0, 1e-6, n, wt, wt, st)
            assert float(out) == 1.25                      # identical models: no term
            comm = torch.empty(n + 2, dtype=dtype, device=DEV)
            l0 = torch.tensor(2.0, device=DEV); l1 = torch.tensor(1.5, device=DEV)
            lib.apts_pack(comm.data_ptr(), wl.data_ptr(), wg.data_ptr(), l0.data_ptr(), l1.data_ptr(), 0, 0.5, 7.0, n, wt, wt, st)
            assert torch.equal(comm[:n], wl - wg) and float(comm[n]) == 0.25 and float(comm[n + 1]) == 7.0
            w = torch.empty_like(wg)
            lib.apts_trial(w.data_ptr(), wg.data_ptr(), comm.data_ptr(), 0, 0.5, 0.0, n, wt, wt, st)
            assert rel_err(w.cpu(), (wg + comm[:n] * 0.5).cpu()) < 1e-7
            lib.apts_trial(w.data_ptr(), wg.data_ptr(), wl.data_ptr(), 1, 1.0, 0.3, n, wt, wt, st)
            assert rel_err(w.cpu(), (wg + (wl - wg).clamp(-0.3, 0.3)).cpu()) < 1e-7
            if dtype == torch.float64:
                # model converted to float64 AFTER construction (trainer.py:1203-1211): the reference's
                # float32 `_step_buffer` / coalesced buffer round the step (apts_d.py:112,196, apts_p.py:108,221)
                c32 = torch.empty(n + 2, dtype=torch.float32, device=DEV)
                lib.apts_pack(c32.data_ptr(), wl.data_ptr(), wg.data_ptr(), l0.data_ptr(), l1.data_ptr(), 0, 0.5, 7.0, n, 1, 0, st)
                assert torch.equal(c32[:n], (wl - wg).float()) and float(c32[n]) == 0.25 and float(c32[n + 1]) == 7.0
                lib.apts_trial(w.data_ptr(), wg.data_ptr(), c32.data_ptr(), 0, 0.5, 0.0, n, 1, 0, st)
                assert torch.equal(w, wg + (c32[:n] * 0.5).double())
                lib.apts_trial(w.data_ptr(), wg.data_ptr(), wl.data_ptr(), 1, 1.0, 0.3, n, 1, 0, st)
                assert torch.equal(w, wg + (wl - wg).float().clamp_(-0.3, 0.3).double())
            # control: accept and reject
            for f1, expect in ((1.0, True), (1.99, False)):
                ds.set(delta=0.1, min_delta=0.01, max_delta=1.0, nu_dec=0.25, nu_inc=0.75, inc_factor=1.2,
                       dec_factor=0.9, tol=1e-6)
                f0t = torch.tensor(2.0, device=DEV); f1t = torch.tensor(f1, device=DEV)
                pred = torch.tensor(1.0, dtype=dtype, device=DEV)
                wcur, w0 = r(), r()
                g0c, G = r(), r()
                w_before, g_before = wcur.clone(), g0c.clone()
                lo = torch.empty((), device=DEV)
                lib.apts_control(ds.sp, ds.wp, f0t.data_ptr(), f1t.data_ptr(), 0, pred.data_ptr(), wt, wcur.data_ptr(),
                                 w0.data_ptr(), g0c.data_ptr(), G.data_ptr(), lo.data_ptr(), None, n, wt, wt, st)
                rep = ds.read()
                assert (rep.accept != 0) == expect
                if expect:
                    assert torch.equal(g0c, G) and torch.equal(wcur, w_before) and float(lo) == f1
                    assert rep.delta == pytest.approx(0.12)
                    assert rep.out_gg == pytest.approx(float(G.double() @ G.double()), rel=1e-9)
                else:
                    assert torch.equal(wcur, w0) and torch.equal(g0c, g_before) and float(lo) == 2.0
                    assert rep.delta == pytest.approx(0.09)


def test_gather_and_seg_copy():
    import ctypes
    lib = _lib.get()
    torch.manual_seed(2)
    ts = [torch.randn(s, device=DEV) for s in (7, 128, 1, 4096, 33)]
    offs = np.concatenate([[0], np.cumsum([t.numel() for t in ts])])
    dst = torch.zeros(int(offs[-1]), device=DEV)
    n = len(ts)
    lib.gather(dst.data_ptr(), n, (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts]),
               (ctypes.c_int64 * n)(*[int(o) for o in offs[:-1]]), (ctypes.c_int64 * n)(*[t.numel() for t in ts]),
               0, _lib.stream())
    assert torch.equal(dst, torch.cat(ts))
    out = torch.full((int(offs[-1]) + 2,), 9.0, device=DEV)
    lib.seg_copy(out.data_ptr(), dst.data_ptr(), 2, (ctypes.c_int64 * 2)(0, 200), (ctypes.c_int64 * 2)(7, 135),
                 (ctypes.c_int64 * 2)(128, 1), out.numel(), 1, 0, _lib.stream())
    exp = torch.zeros_like(out); exp[:128] = dst[7:135]; exp[200] = dst[135]
    assert torch.equal(out, exp)


@pytest.mark.parametrize("n", [1 << 24])
def test_full_size_properties(n):
    """Size-independent properties at a BASELINE-scale vector length."""
    torch.manual_seed(5)
    g = torch.randn(n, device=DEV)
    p, pred = solve_tr_first_order(g, torch.norm(g), 0.25, 1e-6)
    assert float(p.double().norm()) == pytest.approx(0.25, rel=1e-6)            # ||p|| = delta
    assert float(p.double() @ g.double()) == pytest.approx(-float(pred), rel=1e-6)   # g.p = -pred
    h = LSR1(gamma=1.0, memory_length=3, tol=1e-6, device=DEV)
    for i in range(4):
        s = torch.randn(n, device=DEV) * 0.01
        h.update_memory(s, (1.0 + 0.5 * i) * s + 0.001 * torch.randn(n, device=DEV))
    assert len(h._S) == 3
    # linearity of B, symmetry u.Bv = v.Bu
    u, v = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    Bu, Bv = h.B(u).double(), h.B(v).double()
    assert rel_err((h.B(u + 2 * v)).double().cpu(), (Bu + 2 * Bv).cpu()) < 1e-5
    assert float(u.double() @ Bv) == pytest.approx(float(v.double() @ Bu), rel=1e-4, abs=1e-2)
    # B s_i = y_i for every stored pair is NOT guaranteed by compact SR1 with the reference's
    # regularisation, but the secant equation must hold to the regularisation level for the last pair
    S, Y = h.S, h.Y
    assert rel_err(h.B(S[:, -1].contiguous()).cpu(), Y[:, -1].cpu()) < 1e-3


def test_obs_wrapper_and_public_control_step():
    """Reference-named entry points kept for API parity: OBS.solve_tr_subproblem (obs.py:46) on an LSR1
    object, and APTS_Base.control_step(step, pred) (apts_base.py:432)."""
    import torch.nn as nn
    from dd4ml_b200 import APTS_D, OBS
    from dd4ml_b200.workloads import apts_d_yaml_config
    torch.manual_seed(4)
    n = 4096
    h = LSR1(gamma=1.0, memory_length=3, tol=1e-6, device=DEV)
    hc = LSR1Memory(gamma=1.0, memory_length=3, tol=1e-6)
    d = torch.linspace(0.5, 2.0, n)
    for _ in range(3):
        s = torch.randn(n) * 0.01
        h.update_memory(s.to(DEV), (d * s).to(DEV))
        hc.update(s, d * s); hc.precompute()
    g = torch.randn(n)
    p = OBS().solve_tr_subproblem(g.to(DEV), torch.tensor(0.2), h.gamma, h)
    from oracle.obs import obs_solve
    p_or = obs_solve(g, torch.tensor(0.2), hc.gamma, hc.Psi, hc.Minv, eig_sign="canonical")
    assert rel_err(p.cpu(), p_or) < 1e-5
    with pytest.raises(TypeError):
        OBS().solve_tr_subproblem(g.to(DEV), torch.tensor(0.2), h.gamma, torch.zeros(n, 3, device=DEV), None)

    model = nn.Sequential(nn.Flatten(), nn.Linear(16, 4)).to(DEV)
    c = APTS_D.setup_APTS_hparams(apts_d_yaml_config(glob_opt="tr", loc_opt="tr", glob_second_order=False,
                                                     loc_second_order=False, glob_dogleg=False, loc_dogleg=False))
    opt = APTS_D(model.parameters(), model=model, criterion=nn.CrossEntropyLoss(), device=DEV, nr_models=1,
                 glob_opt=c.glob_opt, glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt,
                 loc_opt_hparams=c.loc_opt_hparams, glob_pass=False, foc=False, norm_type=2, max_loc_iters=1,
                 max_glob_iters=1, tol=1e-6, **c.apts_params)
    X, Y = torch.randn(32, 4, 4, device=DEV), torch.randint(0, 4, (32,), device=DEV)
    opt.inputs, opt.labels = X, Y
    opt.init_glob_flat = opt.glob_params_to_vector()
    opt.init_glob_loss = opt.glob_closure_main(compute_grad=True)
    opt.init_glob_grad = opt.glob_grad_to_vector()
    w0 = opt.init_glob_flat.clone()
    step = -0.05 * opt.init_glob_grad / opt.init_glob_grad.norm()
    pred = torch.tensor(0.05 * float(opt.init_glob_grad.norm()), device=DEV)
    loss, grad, delta = opt.control_step(step.clone(), pred)
    assert float(loss) < float(opt.init_glob_loss)                     # a short steepest-descent step is accepted
    assert torch.allclose(opt.glob_params_to_vector(), w0 + step, atol=1e-6)
    loss2, _, delta2 = opt.control_step(-50.0 * step, pred)             # uphill: rejected, parameters restored
    w_acc = w0 + step
    assert delta2 < delta or delta2 == opt.min_delta


def test_memory_length_above_build_maximum_is_refused():
    with pytest.raises(dd4ml_b200.DD4MLError):
        LSR1(memory_length=_lib.get().max_mem_length + 1, device=DEV)


def test_obs_stand_alone_helpers_match_oracle():
    """OBS.phiBar_f / phiBar_fg / Newton / ComputeSBySMW (obs.py:175-254) as
    stand-alone device calls against the same formulas evaluated with torch on the CPU."""
    obs = dd4ml_b200.OBS()
    for dtype, rtol in ((torch.float32, 2e-6), (torch.float64, 1e-12)):
        Lam = torch.tensor([0.5, 1.5, 3.0, 1.0], dtype=dtype)
        a = torch.tensor([0.3, -0.7, 0.2, 0.9], dtype=dtype)
        delta = 0.4
        for sigma in (0.0, 0.7, 5.0):
            D = Lam + sigma
            nrm = torch.norm(a / D)
            f_ref = 1 / nrm - 1 / delta
            g_ref = torch.sum(a ** 2 / D ** 3) / nrm ** 3
            f, gq = obs.phiBar_fg(sigma, Lam.to(DEV), a.to(DEV), delta)
            assert f.dtype == dtype and float(f) == pytest.approx(float(f_ref), rel=rtol)
            assert float(gq) == pytest.approx(float(g_ref), rel=rtol)
            assert float(obs.phiBar_f(sigma, Lam.to(DEV), a.to(DEV), delta)) == pytest.approx(float(f_ref), rel=rtol)
        x = obs.Newton(0.0, Lam.to(DEV), a.to(DEV), delta)
        assert abs(float(obs.phiBar_f(float(x), Lam.to(DEV), a.to(DEV), delta))) <= 1e-6
        assert float(torch.norm(a / (Lam + float(x)))) == pytest.approx(delta, rel=1e-5)
        # degenerate branch (obs.py:242-245): some |a_j| < tol -> (-1/delta, 1/tol), sigma creeps 200 x tol/delta
        a0 = a.clone(); a0[1] = 1e-9
        f, gq = obs.phiBar_fg(0.3, Lam.to(DEV), a0.to(DEV), delta)
        assert float(f) == pytest.approx(-1 / delta, rel=1e-6) and float(gq) == pytest.approx(1e6, rel=1e-6)
        assert float(obs.Newton(0.0, Lam.to(DEV), a0.to(DEV), delta)) == pytest.approx(200 * 1e-6 / delta, rel=1e-3)

    torch.manual_seed(2)
    n, m = 30000, 4
    h = LSR1(gamma=1.0, memory_length=m, tol=1e-6, device=DEV)
    hc = LSR1Memory(gamma=1.0, memory_length=m, tol=1e-6)
    d = torch.linspace(0.5, 2.0, n)
    for _ in range(m):
        s = torch.randn(n) * 0.01
        h.update_memory(s.to(DEV), (d * s).to(DEV))
        assert hc.update(s, d * s)
        hc.precompute()
    g = torch.randn(n)
    for tau in (1.0, 3.7):
        Psi, Minv = hc.Psi.double(), hc.Minv.double()
        W = tau * Minv + Psi.t() @ Psi
        p_ref = -g.double() / tau + Psi @ torch.linalg.solve(W, Psi.t() @ g.double())
        p = obs.ComputeSBySMW(tau, g.to(DEV), None, h, None, None)
        assert rel_err(p.cpu(), p_ref) < 1e-5


def test_obs_explicit_psi_minv_signature_matches_oracle():
    """The reference signature with explicit matrices: OBS.solve_tr_subproblem(g, delta, gamma, Psi, Minv)
    (obs.py:46-173) and OBS.ComputeSBySMW(tau, g, PsiTg, Psi, Minv, PsiPsi) (obs.py:175-182) with (n x k) Psi and
    (k x k) Minv tensors, against the oracle on the same matrices and against the LSR1-object form."""
    from oracle.obs import obs_solve
    torch.manual_seed(5)
    n = 20000
    obs = dd4ml_b200.OBS()
    d = torch.linspace(-0.5, 3.0, n)                       # indefinite: cases A and C both occur
    for m, dtype, tol in ((3, torch.float32, 1e-5), (5, torch.float64, 1e-9)):
        hc = LSR1Memory(gamma=1.0, memory_length=m, tol=1e-6, dtype=dtype)
        h = LSR1(gamma=1.0, memory_length=m, tol=1e-6, device=DEV, dtype=dtype)
        for _ in range(m):
            s = (torch.randn(n) * 0.05).to(dtype)
            y = (d * s + 1e-3 * torch.randn(n)).to(dtype)
            if hc.update(s, y):
                h.update_memory(s.to(DEV), y.to(DEV))
        hc.precompute()
        Psi, Minv, gamma = hc.Psi, hc.Minv, hc.gamma
        g = torch.randn(n).to(dtype)
        for delta in (0.05, 1.0, 50.0):
            info = {}
            p_ref = obs_solve(g, torch.tensor(delta, dtype=dtype), gamma, Psi, Minv, eig_sign="canonical", info=info)
            p = obs.solve_tr_subproblem(g.to(DEV), delta, float(gamma), Psi.to(DEV), Minv.to(DEV))
            assert p.dtype == dtype and rel_err(p.cpu(), p_ref) < tol, (m, delta, info.get("case"))
            assert {1: "A", 3: "C"}[obs.last["case"]] == info["case"]
            p_obj = obs.solve_tr_subproblem(g.to(DEV), delta, float(gamma), h)
            assert rel_err(p.cpu(), p_obj.cpu()) < tol
        tau = 2.5
        W = tau * Minv.double() + Psi.double().t() @ Psi.double()
        s_ref = -g.double() / tau + Psi.double() @ torch.linalg.solve(W, Psi.double().t() @ g.double())
        sp = obs.ComputeSBySMW(tau, g.to(DEV), None, Psi.to(DEV), Minv.to(DEV), None)
        assert rel_err(sp.cpu(), s_ref) < tol
    with pytest.raises(ValueError, match="NaN or Inf"):
        bad = Psi.clone(); bad[0, 0] = float("nan")
        obs.solve_tr_subproblem(g.to(DEV), 1.0, 1.0, bad.to(DEV), Minv.to(DEV))
    with pytest.raises(ValueError, match="non-negative"):
        obs.solve_tr_subproblem(g.to(DEV), -1.0, 1.0, Psi.to(DEV), Minv.to(DEV))
