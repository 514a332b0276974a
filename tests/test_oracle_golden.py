"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py).  CPU only."""
import math

import numpy as np
import pytest
import torch

from _helpers import FFNN, load_golden, make_fg, rel_err, sine_crit

import oracle
from oracle.lsr1 import LSR1Memory
from oracle.subproblem import solve_second_order

torch.set_num_threads(1)

IN_F, HID, OUT = 36, [12, 12], 10
# The oracle runs the same ATen CPU kernels as the reference, so trajectories agree to
# rounding; tolerances leave room for MKL blocking differences (strided vs packed S, Y).
TRAJ_RTOL = 2e-5
VEC_RTOL = 1e-5


def _data(g, dtype=torch.float32):
    return torch.from_numpy(g["X"]).to(dtype), torch.from_numpy(g["Y"])


# --------------------------------------------------------------- kernel level
def test_lsr1_update_precompute_B_match_reference():
    g = load_golden("kernel_level")
    idx = 0
    for m in (1, 3, 5):
        hess = LSR1Memory(gamma=1.0, memory_length=m, tol=1e-6)
        for _ in range(m + 2):
            s, y = torch.from_numpy(g[f"upd{idx}_s"]), torch.from_numpy(g[f"upd{idx}_y"])
            acc = hess.update(s, y)
            assert int(acc) == int(g[f"upd{idx}_accepted"])
            hess.precompute()
            assert float(hess.gamma) == pytest.approx(float(g[f"upd{idx}_gamma"]), rel=1e-6)
            assert rel_err(hess.Minv, g[f"upd{idx}_Minv"]) < 1e-5
            Bv = hess.B(torch.from_numpy(g[f"upd{idx}_v"]))
            assert rel_err(Bv, g[f"upd{idx}_Bv"]) < VEC_RTOL
            idx += 1
    assert idx == int(g["n_updates"])


@pytest.mark.parametrize("m", [1, 3, 5])
@pytest.mark.parametrize("dog", [False, True])
@pytest.mark.parametrize("delta", [0.01, 0.3, 5.0, 100.0])
def test_second_order_solve_matches_reference(m, dog, delta):
    g = load_golden("kernel_level")
    key = f"so_m{m}_dog{int(dog)}_d{delta}"
    S, Y = torch.from_numpy(g[key + "_S"]), torch.from_numpy(g[key + "_Y"])
    hess = LSR1Memory(gamma=float(g[key + "_gamma"]), memory_length=m, tol=1e-6)
    hess.S = [S[:, j].clone() for j in range(S.shape[1])]
    hess.Y = [Y[:, j].clone() for j in range(Y.shape[1])]
    gv = torch.from_numpy(g[key + "_g"])
    info = {}
    p, pred = solve_second_order(gv, torch.norm(gv), delta, hess, 1e-6, dogleg=dog, info=info)
    assert rel_err(p, g[key + "_p"]) < VEC_RTOL
    assert float(pred) == pytest.approx(float(g[key + "_pred"]), rel=1e-4, abs=1e-7)
    assert int(info["case"] == "C") == int(g[key + "_newton"])
    if info["case"] == "C":
        assert info["sigma"] == pytest.approx(float(g[key + "_sigma"]), rel=1e-4, abs=1e-9)


# --------------------------------------------------------------- plain optimizers
def _run_plain(name, make_opt):
    g = load_golden(name)
    X, Y = _data(g)
    model = FFNN(IN_F, HID, OUT)
    fg = make_fg(model, X, Y)
    w = torch.from_numpy(g["w0"]).clone()
    opt = make_opt()
    losses, deltas = [], []
    for i in range(int(g["steps"])):
        loss, _ = opt.step(fg, w)
        losses.append(float(loss))
        deltas.append(float(opt.delta))
        if i == 0:
            assert rel_err(w, g["w_step1"]) < VEC_RTOL
        if i == 4:
            assert rel_err(w, g["w_step5"]) < 1e-4
    np.testing.assert_allclose(losses, g["loss"], rtol=TRAJ_RTOL)
    np.testing.assert_allclose(deltas, g["delta"], rtol=1e-12)
    assert rel_err(w, g["w_final"]) < 1e-3
    return g, opt


@pytest.mark.parametrize("name,so,dog,norm", [
    ("tr_fo", False, False, 2), ("tr_fo_inf", False, False, math.inf),
    ("tr_so", True, False, 2), ("tr_so_dogleg", True, True, 2)])
def test_tr_trajectory_matches_reference(name, so, dog, norm):
    hp = oracle.tr_hparams(0.1, 1e-3, 2.0, norm_type=norm, second_order=so, dogleg=dog)
    g, opt = _run_plain(name, lambda: oracle.TROracle(**hp))
    if so:
        assert len(opt.hess) == int(g["final_pairs"])
        # per-solve vectors recorded from the reference
        solves = [r for r in opt.trace if r.get("k", 0) > 0]
        for i in range(int(g["n_solves"])):
            assert rel_err(solves[i]["step"], g[f"solve{i}_p"]) < 1e-4
            assert solves[i]["pred"] == pytest.approx(float(g[f"solve{i}_pred"]), rel=1e-3, abs=1e-7)


@pytest.mark.parametrize("name,kw", [
    ("lssr1_fo", dict(second_order=False)),
    ("lssr1_so_dog0", dict(second_order=True, dogleg=False)),
    ("lssr1_so_dog1", dict(second_order=True, dogleg=True)),
    ("lssr1_so_dog1_s1", dict(second_order=True, dogleg=True)),
    ("lssr1_so_dog1_s2", dict(second_order=True, dogleg=True)),
    ("lssr1_so_dog1_s3", dict(second_order=True, dogleg=True)),
    ("lssr1_so_dog1_m10", dict(second_order=True, dogleg=True, mem_length=10)),
    ("lssr1_paper", dict(second_order=True, paper_tr_update=True))])
def test_lssr1_trajectory_matches_reference(name, kw):
    kw = dict(kw)
    hp = oracle.lssr1_hparams(0.1, 0.01, 1.0, mem_length=kw.pop("mem_length", 3), **kw)
    hp.pop("sync")
    _run_plain(name, lambda: oracle.LSSR1TROracle(**hp))


# --------------------------------------------------------------- APTS
def _apts_d(name, glob, loc, so, dog):
    g = load_golden(name)
    ws = int(g["world_size"])
    X, Y = _data(g)
    per = X.shape[0] // ws
    models = [FFNN(IN_F, HID, OUT) for _ in range(ws)]
    fgs = [make_fg(models[r], X[r * per:(r + 1) * per], Y[r * per:(r + 1) * per]) for r in range(ws)]
    mk_g = oracle.tr_hparams if glob == "tr" else oracle.lssr1_hparams
    mk_l = oracle.loc_tr_hparams if loc == "tr" else oracle.lssr1_loc_hparams
    ghp = mk_g(0.1, 0.01, 1.0, second_order=so, dogleg=dog)
    lhp = mk_l(0.1, 0.01, 1.0, ws, second_order=so, dogleg=dog)
    opt = oracle.APTSDOracle(torch.from_numpy(g["w0"]), ws, glob, ghp, loc, lhp,
                             **oracle.apts_hparams(0.1, 0.01, 1.0), glob_pass=True, foc=True,
                             norm_type=2, max_loc_iters=3, max_glob_iters=1, tol=1e-6)
    losses, deltas, gev = [], [], []
    for i in range(int(g["steps"])):
        losses.append(float(opt.step(fgs)))
        deltas.append(float(opt.delta))
        gev.append(opt.grad_evals)
        if i == 0:
            assert rel_err(opt.w, g["w_step1"]) < VEC_RTOL
    return g, opt, losses, deltas, gev


@pytest.mark.parametrize("name,glob,loc,so,dog", [
    ("apts_d_tr_fo_ws1", "tr", "tr", False, False),
    ("apts_d_tr_fo_ws2", "tr", "tr", False, False),
    ("apts_d_tr_so_ws2", "tr", "tr", True, False),
    ("apts_d_lssr1_so_ws2", "lssr1_tr", "lssr1_tr", True, True),
    ("apts_d_lssr1_fo_ws4", "lssr1_tr", "lssr1_tr", False, False)])
def test_apts_d_trajectory_matches_reference(name, glob, loc, so, dog):
    g, opt, losses, deltas, gev = _apts_d(name, glob, loc, so, dog)
    np.testing.assert_allclose(losses, g["loss"], rtol=5e-5)
    np.testing.assert_allclose(deltas, g["delta"], rtol=1e-12)
    np.testing.assert_allclose(gev, g["grad_evals"], rtol=0, atol=1e-9)
    assert rel_err(opt.w, g["w_final"]) < 1e-3


@pytest.mark.parametrize("name,glob", [("apts_d_tr_fo_ws8", "tr"), ("apts_d_lssr1_fo_ws8", "lssr1_tr")])
def test_apts_d_eight_subdomains(name, glob):
    """8 gloo ranks of the reference, 16 samples each, 8 outer iterations.  TR: the whole run.  LSSR1_TR: ~130
    strong-Wolfe searches in float32 -- the restatement (subdomains summed in rank order instead of gloo's
    ring order) follows the reference evaluation for evaluation for the first 4 iterations, then one search
    takes a different number of evaluations (iteration 5 here, iteration 6 for the CUDA path): the rule used
    for this case everywhere (tests/test_gpu_multi.py, tools/bench_parity.py) is identity over the first half
    of the run, radii identical and losses within 1e-3 throughout."""
    g, opt, losses, deltas, gev = _apts_d(name, glob, glob, False, False)
    np.testing.assert_allclose(losses, g["loss"], rtol=1e-3)
    np.testing.assert_allclose(deltas, g["delta"], rtol=1e-12)
    same = np.isclose(gev, g["grad_evals"], rtol=0, atol=1e-9)
    if glob == "tr":
        assert same.all()
    else:
        assert (len(same) if same.all() else int(np.argmin(same))) >= len(same) // 2


def _owned_indices(model, tensor_ids):
    sizes = [p.numel() for p in model.parameters()]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return torch.cat([torch.arange(offs[t], offs[t + 1]) for t in sorted(int(t) for t in tensor_ids)])


@pytest.mark.parametrize("name,so,dtype", [
    ("apts_p_tr_fo_ws2", False, torch.float32),
    ("apts_p_tr_fo_ws2_f64", False, torch.float64),
    ("apts_p_tr_so_ws2", True, torch.float32)])
def test_apts_p_trajectory_matches_reference(name, so, dtype):
    g = load_golden(name)
    ws = int(g["world_size"])
    X, Y = _data(g, dtype)
    models = [FFNN(IN_F, HID, OUT).to(dtype) for _ in range(ws)]
    fgs = [make_fg(models[r], X, Y) for r in range(ws)]
    owned = [_owned_indices(models[0], g[f"owned_tensors_rank{r}"]) for r in range(ws)]
    ghp = oracle.tr_hparams(0.1, 0.01, 0.5, norm_type=math.inf, second_order=so)
    lhp = oracle.loc_tr_hparams(0.1, 0.01, 0.5, ws, norm_type=math.inf, second_order=so)
    opt = oracle.APTSPOracle(torch.from_numpy(g["w0"]), owned, "tr", ghp, "tr", lhp,
                             **oracle.apts_hparams(0.1, 0.01, 0.5), glob_pass=True,
                             max_loc_iters=3, max_glob_iters=1, tol=1e-6)
    # in the reference every rank contributes the same full-batch loss; the global
    # closure averages R identical copies
    losses, deltas = [], []
    for _ in range(int(g["steps"])):
        losses.append(float(opt.step(fgs)))
        deltas.append(float(opt.delta))
    np.testing.assert_allclose(losses, g["loss"], rtol=5e-5)
    np.testing.assert_allclose(deltas, g["delta"], rtol=1e-9)
    assert rel_err(opt.w, g["w_final"]) < 1e-3


def test_apts_pinn_trajectory_matches_reference():
    g = load_golden("apts_pinn_ws2")
    ws, overlap = int(g["world_size"]), float(g["overlap"])
    x = torch.from_numpy(g["x"])
    bounds = torch.linspace(0.0, 1.0, ws + 1)
    fgs, fgs_glob, weights = [], [], []
    for r in range(ws):                                   # apts_pinn.py:31-48,84-86
        lo = max(bounds[r].item() - overlap / 2, 0.0)
        hi = min(bounds[r + 1].item() + overlap / 2, 1.0)
        mask = (x[:, 0] >= lo) & (x[:, 0] <= hi)
        xs = x[mask]
        weights.append(int(xs.shape[0]) / x.shape[0])
        ml, mg = FFNN(1, [16, 16], 1, log_softmax=False), FFNN(1, [16, 16], 1, log_softmax=False)
        fgs.append(make_fg(ml, xs, None, crit=sine_crit(xs)))
        fgs_glob.append(make_fg(mg, x, None, crit=sine_crit(x)))
    hp = dict(delta=0.1, nu_dec=0.25, nu_inc=0.75, inc_factor=1.2, dec_factor=0.9,
              max_delta=2.0, min_delta=1e-3, tol=1e-6)
    opt = oracle.APTSDOracle(torch.from_numpy(g["w0"]), ws, "tr", hp, "tr", hp, **{k: v for k, v in hp.items() if k != "tol"},
                             glob_pass=True, foc=True, norm_type=2, max_loc_iters=4,
                             max_glob_iters=1, tol=1e-6)
    losses, deltas = [], []
    for _ in range(int(g["steps"])):
        losses.append(float(opt.step(fgs, weights=weights, fgs_glob=fgs_glob)))
        deltas.append(float(opt.delta))
    np.testing.assert_allclose(losses, g["loss"], rtol=5e-5)
    np.testing.assert_allclose(deltas, g["delta"], rtol=1e-12)
    assert rel_err(opt.w, g["w_final"]) < 1e-3


# --------------------------------------------------------------- TRAdam
@pytest.mark.parametrize("name,norm,lr", [("tradam_inf", math.inf, 0.01), ("tradam_l2", 2, 0.5)])
def test_tradam_trajectory_matches_reference(name, norm, lr):
    g = load_golden(name)
    X, Y = _data(g)
    fg = make_fg(FFNN(IN_F, HID, OUT), X, Y)
    w = torch.from_numpy(g["w0"]).clone()
    opt = oracle.TRAdamOracle(w.numel(), lr=lr, norm_type=norm)
    losses = []
    for i in range(int(g["steps"])):
        losses.append(float(opt.step(fg, w)))
        if i == 0:
            assert rel_err(w, g["w_step1"]) < 1e-7
    np.testing.assert_allclose(losses, g["loss"], rtol=TRAJ_RTOL)
    assert rel_err(w, g["w_final"]) < VEC_RTOL
    assert rel_err(opt.m, g["m_final"]) < VEC_RTOL
    assert rel_err(opt.v, g["v_final"]) < VEC_RTOL
    assert rel_err(opt.step_buf, g["step_final"]) < VEC_RTOL


# --------------------------------------------------------------- how much of the reference is LAPACK-sign defined
def test_sign_sensitivity_census_of_the_reference_goldens():
    """OBS starts Newton from max_j(a_j/delta - Lambda_j) with SIGNED a_j (obs.py:162-168), i.e. from the signs
    `torch.linalg.eigh` happens to give the eigenvectors.  Census over every sub-problem solve the goldens
    record (24 kernel-level + 40 inside the TR / LSSR1_TR trajectories): a solve is sign-sensitive when sigma*
    differs between sign patterns.  Measured: 0 of 24 and 1 of 40 -- `lssr1_so_dog1` solve 6 (spread 0.19 of
    sigma*; the step changes by 99 %), which is where that trajectory leaves the canonical-sign one (step 7).
    On the other 63 the oracle reproduces the reference's step to <= 3e-7 in EITHER sign mode.  The same
    census states the reference's own float32 noise: its step differs from the float64 evaluation of the same
    algorithm by up to 7e-3 on ill-conditioned memories."""
    from _helpers import golden_solves, oracle_memory, sigma_spread
    from oracle.subproblem import solve_second_order as oso
    torch.set_num_threads(1)
    sensitive, worst_insensitive, ref_noise = [], 0.0, 0.0
    n = 0
    for name in ("tr_so", "tr_so_dogleg", "lssr1_so_dog0", "lssr1_so_dog1", "lssr1_paper"):
        for i, r in enumerate(golden_solves(name)):
            g = torch.from_numpy(r["g"])
            delta, dog, tol = float(r["delta"]), bool(int(r["dogleg"])), float(r["tol"])
            info = {}
            p_can, _ = oso(g, torch.norm(g), delta, oracle_memory(r["S"], r["Y"], r["gamma"]), tol, dogleg=dog,
                           eig_sign="canonical", info=info)
            p64, _ = oso(g.double(), torch.norm(g.double()), delta,
                         oracle_memory(r["S"], r["Y"], r["gamma"], torch.float64), tol, dogleg=dog, eig_sign="canonical")
            n += 1
            if sigma_spread(info, delta, float(r["gamma"])) > 1e-6:
                sensitive.append((name, i))
            else:
                worst_insensitive = max(worst_insensitive, rel_err(p_can, r["p"]))
                ref_noise = max(ref_noise, rel_err(r["p"], p64))
    assert n == 40 and sensitive == [("lssr1_so_dog1", 6)]
    assert worst_insensitive < 1e-6
    assert 1e-3 < ref_noise < 2e-2          # the reference's float32 pipeline vs its float64 evaluation
