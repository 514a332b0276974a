"""GPU parity on the model families of BASELINE.json's other configs (the bench line is the FFNN one):
C1 DeepONet + plain TR, C3 CNN + APTS_D with the L-SR1 second-order sub-problem, C4 minGPT-style
transformer + APTS_D / LSSR1_TR.  The models are small independent re-statements of the
architectures (reference: models/deeponet/deeponet.py:9-26, models/cnn/simple_cnn.py:81-103,
models/gpt/base_gpt.py) -- what is under test is the optimizer path on their parameter layouts
(4-D conv kernels, embeddings, LayerNorm, 77-tensor lists), against the CPU oracle on the same
seeded inputs.

Tolerances: first-order runs have no eigen-solver and few discrete decisions -> whole loss
trajectory within 1e-3 relative, radius sequence identical; second-order runs -> first outer
iteration within 1e-4 and the trajectory within 5e-3 of one of the oracle's branches under 1e-6
perturbations (tests/_helpers.py::branch_distance).  cuDNN's TF32 convolutions are switched off:
the reference path is fp32.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from _helpers import branch_distance, noisy_fg, rel_err, set_flat

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from dd4ml_b200 import APTS_D, LSSR1_TR, TR
    from dd4ml_b200.utility.optimizer_utils import get_lssr1_tr_hparams, get_tr_hparams

import oracle

DEV = "cuda"


class Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def cfg(**kw):
    base = dict(delta=0.1, min_delta=0.01, max_delta=1.0, norm_type=2, glob_second_order=False,
                glob_dogleg=False, loc_second_order=False, loc_dogleg=False, mem_length=3,
                max_wolfe_iters=5, max_zoom_iters=5, tol=1e-6, paper_tr_update=False)
    base.update(kw)
    return Cfg(**base)


# ------------------------------------------------------------------------------ models
class DeepONet(nn.Module):
    """branch 50->64->64, trunk 1->64->64, output = <branch, trunk> (n = 11 712)."""

    def __init__(self, sensors=50, width=64):
        super().__init__()
        self.branch = nn.Sequential(nn.Linear(sensors, width), nn.Tanh(), nn.Linear(width, width))
        self.trunk = nn.Sequential(nn.Linear(1, width), nn.Tanh(), nn.Linear(width, width))

    def forward(self, xb):
        u, y = xb
        return (self.branch(u) * self.trunk(y)).sum(dim=1, keepdim=True)


class SmallCNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3, 8, 3, padding=1)
        self.c2 = nn.Conv2d(8, 16, 3, padding=1)
        self.f1 = nn.Linear(16 * 8 * 8, 32)
        self.f2 = nn.Linear(32, 10)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.c1(x)), 2)
        x = F.max_pool2d(F.relu(self.c2(x)), 2)
        return self.f2(F.relu(self.f1(x.flatten(1))))


class Block(nn.Module):
    def __init__(self, d, h, T):
        super().__init__()
        self.ln1, self.ln2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.qkv, self.proj = nn.Linear(d, 3 * d), nn.Linear(d, d)
        self.fc, self.out = nn.Linear(d, 4 * d), nn.Linear(4 * d, d)
        self.h = h
        self.register_buffer("mask", torch.tril(torch.ones(T, T)).view(1, 1, T, T))

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = self.qkv(self.ln1(x)).split(C, dim=2)
        q, k, v = (t.view(B, T, self.h, C // self.h).transpose(1, 2) for t in (q, k, v))
        att = (q @ k.transpose(-2, -1)) / math.sqrt(C // self.h)
        att = att.masked_fill(self.mask[:, :, :T, :T] == 0, float("-inf")).softmax(dim=-1)
        x = x + self.proj((att @ v).transpose(1, 2).reshape(B, T, C))
        return x + self.out(F.gelu(self.fc(self.ln2(x))))


class TinyGPT(nn.Module):
    """char-level decoder (vocab 65 as tinyshakespeare), no dropout (parity needs determinism)."""

    def __init__(self, vocab=65, T=32, d=48, h=3, layers=2):
        super().__init__()
        self.tok, self.pos = nn.Embedding(vocab, d), nn.Embedding(T, d)
        self.blocks = nn.ModuleList([Block(d, h, T) for _ in range(layers)])
        self.ln, self.head = nn.LayerNorm(d), nn.Linear(d, vocab, bias=False)

    def forward(self, idx):
        x = self.tok(idx) + self.pos(torch.arange(idx.shape[1], device=idx.device))[None]
        for b in self.blocks:
            x = b(x)
        return self.head(self.ln(x))


def token_ce(out, y):
    return F.cross_entropy(out.reshape(-1, out.shape[-1]), y.reshape(-1).long())


def flat_of(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()


def cpu_fg(model, X, Y, crit):
    def fg(w):
        set_flat(model, w)
        for p in model.parameters():
            p.grad = None
        loss = crit(model(X), Y)
        loss.backward()
        return loss.detach(), torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    return fg


def to_dev(x):
    return tuple(t.to(DEV) for t in x) if isinstance(x, tuple) else x.to(DEV)


@pytest.fixture(autouse=True)
def _fp32_convolutions():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


# ------------------------------------------------------------------------------ C1: DeepONet + TR
def _deeponet_data(seed=0, nfun=20, npts=50):
    gen = torch.Generator().manual_seed(seed)
    a = torch.rand(nfun, 1, generator=gen)
    xs = torch.linspace(0, 1, 50)[None]
    u = a * torch.sin(2 * math.pi * xs)                     # sensor values of a*sin(2 pi x)
    y = torch.rand(nfun, npts, 1, generator=gen)
    target = a[:, None] * (1 - torch.cos(2 * math.pi * y)) / (2 * math.pi)   # antiderivative
    U = u[:, None, :].expand(nfun, npts, 50).reshape(-1, 50).contiguous()
    return (U, y.reshape(-1, 1)), target.reshape(-1, 1)


@pytest.mark.parametrize("so,dog", [(False, False), (True, False), (True, True)])
def test_c1_deeponet_plain_tr(so, dog):
    torch.manual_seed(42)
    X, Y = _deeponet_data()
    ref = DeepONet()
    assert sum(p.numel() for p in ref.parameters()) == 11712
    w0 = flat_of(ref).clone()
    model = DeepONet().to(DEV)
    set_flat(model, w0.to(DEV))
    hp = get_tr_hparams(cfg(delta=0.1, min_delta=1e-3, max_delta=2.0, glob_second_order=so, glob_dogleg=dog))
    assert hp["mem_length"] == 5          # optimizer_utils.py:123 hard-codes it: KMAX = 6 kernels
    opt = TR(model.parameters(), **hp)
    Xd, Yd = to_dev(X), Y.to(DEV)
    crit = nn.MSELoss()

    def closure(compute_grad=True):
        opt.zero_grad()
        loss = crit(model(Xd), Yd)
        if compute_grad:
            loss.backward()
        return loss

    ohp = oracle.tr_hparams(0.1, 1e-3, 2.0, norm_type=2, second_order=so, dogleg=dog)
    ofg = cpu_fg(DeepONet(), X, Y, crit)
    oopt = oracle.TROracle(**ohp, eig_sign="canonical")
    ow = w0.clone()
    losses, olosses, deltas, odeltas = [], [], [], []
    # delta = 0.1 overshoots on this problem: ~14 rejections (undo path) before the first accept
    for i in range(40):
        loss, _ = opt.step(closure)
        oloss, _ = oopt.step(ofg, ow)
        losses.append(float(loss.detach())); olosses.append(float(oloss))
        deltas.append(float(opt.delta)); odeltas.append(float(oopt.delta))
    print(f"deeponet TR so={so} dogleg={dog}: {losses[0]:.6f} -> {losses[-1]:.6f} (oracle {olosses[-1]:.6f})")
    assert losses[-1] < losses[0]
    if not so:
        np.testing.assert_allclose(losses, olosses, rtol=1e-3)
        np.testing.assert_allclose(deltas, odeltas, rtol=1e-9)
        assert rel_err(flat_of(model), ow) < 1e-4
    else:
        np.testing.assert_allclose(deltas[:16], odeltas[:16], rtol=1e-9)   # up to the first accepted steps

        def orun(eps, seed):
            gen = torch.Generator().manual_seed(seed)
            o = oracle.TROracle(**ohp, eig_sign="canonical")
            fg = noisy_fg(cpu_fg(DeepONet(), X, Y, crit), eps, gen)
            w = w0.clone()
            return [float(o.step(fg, w)[0]) for _ in range(24)]
        d, br = branch_distance(losses, orun, 22, seeds=12)
        print(f"  distance to the closest oracle branch over 22 steps: {d:.2e}")
        assert d < 5e-3, (losses[:22], br)


# ------------------------------------------------------------------------------ C3 / C4: APTS_D
def _apts_d(model, crit, glob, loc, c):
    c.glob_opt, c.loc_opt = glob, loc
    c = APTS_D.setup_APTS_hparams(c)
    return APTS_D(model.parameters(), model=model, criterion=crit, device=DEV, nr_models=1,
                  glob_opt=c.glob_opt, glob_opt_hparams=c.glob_opt_hparams, loc_opt=c.loc_opt,
                  loc_opt_hparams=c.loc_opt_hparams, glob_pass=True, norm_type=c.norm_type, max_loc_iters=3,
                  max_glob_iters=1, tol=1e-6, foc=True, **c.apts_params)


def _oracle_apts_d(w0, glob, loc, so, dog, mem):
    mk_g = oracle.tr_hparams if glob == "tr" else oracle.lssr1_hparams
    mk_l = oracle.loc_tr_hparams if loc == "tr" else oracle.lssr1_loc_hparams
    kw = dict(second_order=so, dogleg=dog)
    ghp, lhp = mk_g(0.1, 0.01, 1.0, **kw), mk_l(0.1, 0.01, 1.0, 1, **kw)
    if glob != "tr":
        ghp["mem_length"] = mem
    if loc != "tr":
        lhp["mem_length"] = mem
    return oracle.APTSDOracle(w0.clone(), 1, glob, ghp, loc, lhp, **oracle.apts_hparams(0.1, 0.01, 1.0),
                              glob_pass=True, foc=True, norm_type=2, max_loc_iters=3, max_glob_iters=1,
                              tol=1e-6, eig_sign="canonical")


def _run_apts_d(make_model, X, Y, crit, glob, loc, so, dog, iters, branch_steps=3):
    torch.manual_seed(7)
    w0 = flat_of(make_model()).clone()
    model = make_model().to(DEV)
    set_flat(model, w0.to(DEV))
    opt = _apts_d(model, crit, glob, loc, cfg(glob_second_order=so, glob_dogleg=dog, loc_second_order=so,
                                              loc_dogleg=dog, mem_length=3))
    oopt = _oracle_apts_d(w0, glob, loc, so, dog, 3)
    ofg = [cpu_fg(make_model(), X, Y, crit)]
    Xd, Yd = X.to(DEV), Y.to(DEV)
    losses, olosses, deltas, odeltas = [], [], [], []
    for _ in range(iters):
        losses.append(float(opt.step(Xd, Yd)))
        olosses.append(float(oopt.step(ofg)))
        deltas.append(float(opt.delta)); odeltas.append(float(oopt.delta))
    assert losses[-1] < losses[0]
    assert abs(losses[0] - olosses[0]) <= 1e-4 * abs(olosses[0])
    if not so:
        np.testing.assert_allclose(losses, olosses, rtol=1e-3)
        np.testing.assert_allclose(deltas, odeltas, rtol=1e-9)
    else:
        def orun(eps, seed):
            gen = torch.Generator().manual_seed(seed)
            o = _oracle_apts_d(w0, glob, loc, so, dog, 3)
            fgs = [noisy_fg(cpu_fg(make_model(), X, Y, crit), eps, gen)]
            return [float(o.step(fgs)) for _ in range(4)]
        d, br = branch_distance(losses, orun, branch_steps, seeds=10)
        print(f"  distance to the closest oracle branch over {branch_steps} outer steps: {d:.2e}")
        assert d < 5e-3, (losses[:branch_steps], br)
    return losses, olosses


@pytest.mark.parametrize("glob,loc,so,dog", [("tr", "tr", False, False), ("tr", "tr", True, False),
                                             ("lssr1_tr", "lssr1_tr", True, True)])
def test_c3_cnn_apts_d(glob, loc, so, dog):
    gen = torch.Generator().manual_seed(3)
    X = torch.randn(48, 3, 32, 32, generator=gen)
    Y = torch.randint(0, 10, (48,), generator=gen)
    losses, olosses = _run_apts_d(SmallCNN, X, Y, nn.CrossEntropyLoss(), glob, loc, so, dog, 5)
    print(f"cnn apts_d {glob}/{loc} so={so}: {losses[0]:.5f} -> {losses[-1]:.5f} (oracle {olosses[-1]:.5f})")


@pytest.mark.parametrize("glob,loc,so,dog", [("tr", "tr", False, False), ("lssr1_tr", "lssr1_tr", True, True)])
def test_c4_gpt_apts_d(glob, loc, so, dog):
    gen = torch.Generator().manual_seed(5)
    tokens = torch.randint(0, 65, (8, 33), generator=gen)
    X, Y = tokens[:, :-1].contiguous(), tokens[:, 1:].contiguous()
    losses, olosses = _run_apts_d(TinyGPT, X, Y, token_ce, glob, loc, so, dog, 5)
    print(f"gpt apts_d {glob}/{loc} so={so}: {losses[0]:.5f} -> {losses[-1]:.5f} (oracle {olosses[-1]:.5f})")


def test_c4_gpt_lssr1_tr_top_level():
    """LSSR1_TR as the top-level optimizer on the transformer (`optimizer: lssr1_tr`)."""
    gen = torch.Generator().manual_seed(5)
    tokens = torch.randint(0, 65, (8, 33), generator=gen)
    X, Y = tokens[:, :-1].contiguous(), tokens[:, 1:].contiguous()
    torch.manual_seed(11)
    w0 = flat_of(TinyGPT()).clone()
    model = TinyGPT().to(DEV)
    set_flat(model, w0.to(DEV))
    hp = get_lssr1_tr_hparams(cfg(glob_second_order=False))
    hp["sync"] = False
    opt = LSSR1_TR(model.parameters(), **hp)
    ohp = oracle.lssr1_hparams(0.1, 0.01, 1.0, mem_length=3, second_order=False, dogleg=False)
    ohp.pop("sync")
    Xd, Yd = X.to(DEV), Y.to(DEV)

    def closure(compute_grad=True):
        opt.zero_grad()
        loss = token_ce(model(Xd), Yd)
        if compute_grad:
            loss.backward()
        return loss

    oopt = oracle.LSSR1TROracle(**ohp, eig_sign="canonical")
    ofg = cpu_fg(TinyGPT(), X, Y, token_ce)
    ow = w0.clone()
    losses, olosses = [], []
    for _ in range(8):
        losses.append(float(opt.step(closure)[0].detach()))
        olosses.append(float(oopt.step(ofg, ow)[0]))
    print(f"gpt LSSR1_TR first-order: {losses[0]:.5f} -> {losses[-1]:.5f} (oracle {olosses[-1]:.5f})")
    np.testing.assert_allclose(losses, olosses, rtol=1e-3)
    assert rel_err(flat_of(model), ow) < 5e-3


# ------------------------------------------------------------------------------ C3 / C4 at their NATIVE sizes
def _native_cnn():
    from dd4ml_b200.workloads import SimpleCNN
    return SimpleCNN(p_drop=0.0)          # n = 620 810; Dropout off so that GPU and CPU evaluate the same function


def _native_gpt():
    from dd4ml_b200.workloads import MiniGPT
    return MiniGPT(p_drop=0.0)            # gpt-mini, n = 2 719 104 (> 2^20: every vector pass takes the TMA kernels)


@pytest.mark.parametrize("glob,loc,so,dog", [("tr", "tr", False, False), ("lssr1_tr", "lssr1_tr", True, True)])
def test_c3_cnn_native_size_apts_d(glob, loc, so, dog):
    """BASELINE config 3 at its real size: `simple_cnn` (models/cnn/simple_cnn.py:81-103, n = 620 810) on a
    CIFAR-10-shape batch, APTS_D against the oracle.  BatchNorm in training mode (batch statistics)."""
    from dd4ml_b200.workloads import cifar_shape_batch
    assert sum(p.numel() for p in _native_cnn().parameters()) == 620810
    X, Y = cifar_shape_batch(64, 11)
    # second order: cuDNN's float32 convolution gradients differ from the CPU's by ~1e-5 (different summation
    # order over 64 x 32 x 32 positions), ten times the perturbation the oracle ensemble is built with, and the
    # loss falls by a factor 2 per outer step here: the first two outer steps (8 inner LSSR1_TR steps, ~18
    # evaluations) must sit on an oracle branch, the later ones only have to keep descending
    losses, olosses = _run_apts_d(_native_cnn, X, Y, nn.CrossEntropyLoss(), glob, loc, so, dog, 4,
                                  branch_steps=2 if so else 3)
    print(f"native cnn apts_d {glob}/{loc} so={so}: {losses[0]:.5f} -> {losses[-1]:.5f} (oracle {olosses[-1]:.5f})")


@pytest.mark.parametrize("glob,loc,so,dog", [("tr", "tr", False, False), ("lssr1_tr", "lssr1_tr", True, True)])
def test_c4_gpt_mini_native_size_apts_d(glob, loc, so, dog):
    """BASELINE config 4 at its real size: minGPT `gpt-mini` (models/gpt/base_gpt.py:181, n = 2 719 104, the only
    config above the 2^20 TMA threshold) on 8 x 128 synthetic tokens, dropout 0, APTS_D against the oracle."""
    from dd4ml_b200.workloads import token_batch, token_cross_entropy
    assert sum(p.numel() for p in _native_gpt().parameters()) == 2719104
    X, Y = token_batch(8, 13)
    losses, olosses = _run_apts_d(_native_gpt, X, Y, token_cross_entropy, glob, loc, so, dog, 3)
    print(f"native gpt-mini apts_d {glob}/{loc} so={so}: {losses[0]:.5f} -> {losses[-1]:.5f} (oracle {olosses[-1]:.5f})")
