"""Synthetic workloads of the BASELINE configs (bench / smoke support).

Models are plain PyTorch re-statements of the reference architectures (random init of
that architecture, synthetic data of that shape -- there is no network for datasets);
their forward/backward is untouched PyTorch, as the north star requires.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class MediumFFNN(nn.Module):
    """784 -> [128] x 8 -> 10 with log_softmax (reference models/ffnn/medium_ffnn.py:6-39,
    tests/config_files/config_apts_d.yaml:48-49): n = 217 354 parameters, 18 tensors."""

    def __init__(self, in_features=784, hidden=(128,) * 8, classes=10):
        super().__init__()
        dims = [in_features, *hidden]
        self.stages = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.finish = nn.Linear(dims[-1], classes)

    def forward(self, x):
        x = x.reshape(x.size(0), -1)
        for lin in self.stages:
            x = F.relu(lin(x))
        return F.log_softmax(self.finish(x), dim=1)


def mnist_shape_batch(n_samples, seed, device="cpu", pin=False):
    """x ~ N(0,1) of shape (B, 1, 28, 28), y ~ U{0..9}."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_samples, 1, 28, 28, generator=g)
    y = torch.randint(0, 10, (n_samples,), generator=g)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return x.to(device), y.to(device)


class Cfg:
    """Attribute bag with the keys of the reference's W&B/yaml config contract."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def apts_d_yaml_config(**over):
    """tests/config_files/config_apts_d.yaml:6-83 (the configuration BASELINE.json's
    configs[1] is quoted on)."""
    c = dict(delta=0.1, min_delta=0.01, max_delta=1.0, norm_type=2, glob_second_order=True,
             glob_dogleg=True, loc_second_order=True, loc_dogleg=True, mem_length=3, max_wolfe_iters=5,
             max_zoom_iters=5, tol=1e-6, paper_tr_update=False, glob_opt="lssr1_tr", loc_opt="lssr1_tr",
             max_loc_iters=3, max_glob_iters=1, glob_pass=True, foc=True)
    c.update(over)
    return Cfg(**c)


class PINNNet(nn.Module):
    """1 -> [20] x 8 -> 1 with tanh (reference models/ffnn/pinn_ffnn.py:11-29): n = 3 001 parameters,
    18 tensors in the same order (weight, bias per layer)."""

    def __init__(self, in_features=1, hidden=(20,) * 8, out_features=1):
        super().__init__()
        dims = [in_features, *hidden]
        mods = []
        for a, b in zip(dims[:-1], dims[1:]):
            mods += [nn.Linear(a, b), nn.Tanh()]
        mods.append(nn.Linear(dims[-1], out_features))
        self.net = nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)


class AllenCahnLoss(nn.Module):
    """Residual loss of the 1-D Allen-Cahn PINN, -u'' + u^3 - u = 0 on [low, high], u(low) = 1,
    u(high) = -1 (reference utility/pinn_allencahn_loss.py:14-60): mean over ALL points of the squared
    interior residual masked by (1 - flag) plus mean over all points of the squared boundary error
    masked by flag.  ``current_x`` is the collocation tensor the prediction was computed from (set by
    the caller: trainer.py:997-998, apts_pinn.py:_set_criterion_input)."""

    def __init__(self, current_x=None, low=0.0, high=1.0):
        super().__init__()
        self.current_x, self.low, self.high = current_x, low, high

    def forward(self, u, flag):
        x = self.current_x
        if x is None:
            raise ValueError("current_x must be set before calling AllenCahnLoss")
        if not x.requires_grad:
            x.requires_grad_(True)
        ones = torch.ones_like(u)
        du, = torch.autograd.grad(u, x, ones, create_graph=True, retain_graph=True)
        d2u, = torch.autograd.grad(du, x, torch.ones_like(du), create_graph=True)
        flag = flag.squeeze()
        xs, us = x.squeeze(), u.squeeze()
        res = (-d2u + u ** 3 - u).squeeze()
        target = torch.zeros_like(xs)
        target[torch.isclose(xs, torch.tensor(self.low, device=x.device, dtype=x.dtype))] = 1.0
        target[torch.isclose(xs, torch.tensor(self.high, device=x.device, dtype=x.dtype))] = -1.0
        return (res ** 2 * (1 - flag)).mean() + ((us - target) ** 2 * flag).mean()


def allen_cahn_points(n_interior=1000, low=0.0, high=1.0):
    """Collocation set of datasets/pinn_allencahn.py:38-64: n_interior equispaced interior points plus
    the two boundary points, sorted; returns (x [N,1] fp32, boundary flag [N,1] fp32)."""
    step = (high - low) / (n_interior + 1)
    interior = torch.linspace(low + step, high - step, n_interior, dtype=torch.float32).unsqueeze(1)
    boundary = torch.tensor([[low], [high]], dtype=torch.float32)
    data = torch.cat([boundary, interior], 0)
    mask = torch.cat([torch.ones(2, 1), torch.zeros(n_interior, 1)], 0)
    idx = torch.argsort(data[:, 0])
    return data[idx], mask[idx]


# ---------------------------------------------------------------------------------------------------
# The other BASELINE models at their native sizes (forward/backward = plain PyTorch, untouched)
class DeepONet(nn.Module):
    """Branch 50 -> 64 -> 64, trunk 1 -> 64 -> 64, ReLU after every layer, y = sum(b * t)
    (reference models/deeponet/deeponet.py:9-42): n = 11 712 parameters, 8 tensors."""

    def __init__(self, sensors=50, width=(64, 64), trunk_in=1):
        super().__init__()

        def mlp(i):
            mods = []
            for h in width:
                mods += [nn.Linear(i, h), nn.ReLU()]
                i = h
            return nn.Sequential(*mods)
        self.branch_net, self.trunk_net = mlp(sensors), mlp(trunk_in)

    def forward(self, inputs):
        branch, trunk = inputs
        return (self.branch_net(branch) * self.trunk_net(trunk)).sum(dim=1, keepdim=True)


def deeponet_sine_batch(n_items=1000, seed=42, n_samples=1000, n_sensors=50, n_trunk=100):
    """First `n_items` items of the reference's SineOperatorDataset (datasets/deeponet_sine.py:9-45):
    functions a -> sin(a x), a ~ U(0,1), sampled at 50 sensors / 100 trunk points on [0, 2 pi];
    returns (branch [B,50], trunk [B,1], y [B,1]) with float targets."""
    import math
    g = torch.Generator().manual_seed(seed)
    hi = 2 * math.pi
    sensors, trunk = torch.linspace(0.0, hi, n_sensors), torch.linspace(0.0, hi, n_trunk)
    a = torch.rand(n_samples, generator=g)
    idx = torch.arange(n_items)
    s_idx, t_idx = idx // n_trunk, idx % n_trunk
    branch = torch.sin(a[s_idx].unsqueeze(1) * sensors)
    x = trunk[t_idx].unsqueeze(1)
    return branch, x, torch.sin(a[s_idx].unsqueeze(1) * x)


class SimpleCNN(nn.Module):
    """3 x (conv3x3 -> BatchNorm -> ReLU -> maxpool2) with 32/64/128 channels, fc 2048 -> 256 -> 10,
    Dropout(0.5) (reference models/cnn/simple_cnn.py:81-110) on CIFAR-10-shape input:
    n = 620 810 parameters, 16 tensors."""

    def __init__(self, in_ch=3, hw=32, classes=10, p_drop=0.5):
        super().__init__()
        chans = [in_ch, 32, 64, 128]
        self.convs = nn.ModuleList([nn.Conv2d(a, b, 3, padding=1) for a, b in zip(chans[:-1], chans[1:])])
        self.norms = nn.ModuleList([nn.BatchNorm2d(c) for c in chans[1:]])
        self.fc1 = nn.Linear(chans[-1] * (hw // 8) ** 2, 256)
        self.drop = nn.Dropout(p_drop)
        self.fc2 = nn.Linear(256, classes)

    def forward(self, x):
        for conv, bn in zip(self.convs, self.norms):
            x = F.max_pool2d(F.relu(bn(conv(x))), 2)
        x = F.relu(self.fc1(x.flatten(1)))
        return self.fc2(self.drop(x))

    def parameters_in_reference_order(self):
        """conv1, bn1, conv2, bn2, conv3, bn3, fc1, fc2 -- the registration order of the reference
        class (its `parameters()` order), for loading a flat reference vector."""
        out = []
        for conv, bn in zip(self.convs, self.norms):
            out += list(conv.parameters()) + list(bn.parameters())
        return out + list(self.fc1.parameters()) + list(self.fc2.parameters())


def cifar_shape_batch(n_samples, seed, pin=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_samples, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (n_samples,), generator=g)
    return (x.pin_memory(), y.pin_memory()) if pin else (x, y)


class _GPTBlock(nn.Module):
    def __init__(self, d, heads, block, p):
        super().__init__()
        self.ln_1, self.ln_2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.c_attn, self.c_proj = nn.Linear(d, 3 * d), nn.Linear(d, d)
        self.c_fc, self.c_proj2 = nn.Linear(d, 4 * d), nn.Linear(4 * d, d)
        self.heads = heads
        self.attn_drop, self.resid_drop, self.mlp_drop = nn.Dropout(p), nn.Dropout(p), nn.Dropout(p)
        self.register_buffer("mask", torch.tril(torch.ones(block, block)).view(1, 1, block, block))

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = self.c_attn(self.ln_1(x)).split(C, dim=2)
        sh = lambda t: t.view(B, T, self.heads, C // self.heads).transpose(1, 2)  # noqa: E731
        q, k, v = sh(q), sh(k), sh(v)
        att = (q @ k.transpose(-2, -1)) * (1.0 / (k.size(-1) ** 0.5))
        att = att.masked_fill(self.mask[:, :, :T, :T] == 0, float("-inf"))
        att = self.attn_drop(F.softmax(att, dim=-1))
        y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
        x = x + self.resid_drop(self.c_proj(y))
        h = self.c_fc(self.ln_2(x))
        h = 0.5 * h * (1.0 + torch.tanh(0.7978845608028654 * (h + 0.044715 * torch.pow(h, 3.0))))
        return x + self.mlp_drop(self.c_proj2(h))


class MiniGPT(nn.Module):
    """minGPT `gpt-mini`: 6 layers, 6 heads, width 192, vocab 65, block 128, dropout 0.1
    (reference models/gpt/base_gpt.py:181, models/gpt/nanogpt/model.py:31-92): n = 2 719 104
    parameters, 77 tensors (wte, wpe, 6 x 12, ln_f x 2, lm_head without bias)."""

    def __init__(self, vocab=65, block=128, layers=6, heads=6, d=192, p_drop=0.1):
        super().__init__()
        self.wte, self.wpe = nn.Embedding(vocab, d), nn.Embedding(block, d)
        self.h = nn.ModuleList([_GPTBlock(d, heads, block, p_drop) for _ in range(layers)])
        self.ln_f = nn.LayerNorm(d)
        self.lm_head = nn.Linear(d, vocab, bias=False)
        self.drop = nn.Dropout(p_drop)
        for m in self.modules():                      # GPT-2 initialisation (model.py:83-92, 49-54)
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, mean=0.0, std=0.02)
                if isinstance(m, nn.Linear) and m.bias is not None:
                    nn.init.zeros_(m.bias)
        for blk in self.h:
            nn.init.normal_(blk.c_proj.weight, mean=0.0, std=0.02 / (2 * layers) ** 0.5)
            nn.init.normal_(blk.c_proj2.weight, mean=0.0, std=0.02 / (2 * layers) ** 0.5)

    def forward(self, idx):
        T = idx.size(1)
        pos = torch.arange(T, device=idx.device).unsqueeze(0)
        x = self.drop(self.wte(idx) + self.wpe(pos))
        for blk in self.h:
            x = blk(x)
        return self.lm_head(self.ln_f(x))


def token_cross_entropy(logits, targets):
    """utility/ml_utils.py:33-36 (`cross_entropy_transformers`)."""
    return F.cross_entropy(logits.view(-1, logits.size(-1)), targets.view(-1), ignore_index=-1)


def token_batch(batch, seed, block=128, vocab=65, pin=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, vocab, (batch, block), generator=g)
    y = torch.randint(0, vocab, (batch, block), generator=g)
    return (x.pin_memory(), y.pin_memory()) if pin else (x, y)


def teacher_labels(X, seed=7, hidden=64, classes=10):
    """Labels a fixed random teacher MLP assigns to X (learnable targets: with uniform random labels
    the N >= 2 APTS_D runs stall at ln 10 after one outer iteration, which makes per-N work unequal)."""
    g = torch.Generator().manual_seed(seed)
    d = X[0].numel()
    w1 = torch.randn(d, hidden, generator=g) / d ** 0.5
    w2 = torch.randn(hidden, classes, generator=g) / hidden ** 0.5
    return (torch.tanh(X.reshape(X.size(0), -1).float() @ w1) @ w2).argmax(dim=1)
