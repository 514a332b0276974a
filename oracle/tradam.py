"""TEST INFRASTRUCTURE ONLY -- CPU restatement of TRAdam.step on a flat vector.

Follows /root/reference/src/dd4ml/optimizers/tradam.py:80-111 line by line (same torch ops, so
the fp32 rounding sequence is the reference's).  Pinned by tests/golden/tradam_*.npz, generated
from the unmodified reference class (tests/golden/gen_golden.py).  Never imported by the product.
"""
from __future__ import annotations

import torch


class TRAdamOracle:
    def __init__(self, n, *, lr, betas=(0.9, 0.999), eps=1e-8, norm_type=torch.inf, dtype=torch.float32):
        if norm_type not in (2, torch.inf):
            raise ValueError("norm_type must be 2 or torch.inf")       # tradam.py:23-24
        self.lr, self.betas, self.eps, self.norm_type = float(lr), betas, float(eps), norm_type
        self.m = torch.zeros(n, dtype=dtype)                             # tradam.py:47-48
        self.v = torch.zeros(n, dtype=dtype)
        self.step_buf = torch.zeros(n, dtype=dtype)
        self.t = 0
        self.step_len = None

    def step(self, fg, w):
        """fg(w) -> (loss, grad); updates w in place; returns the loss at the incoming w."""
        self.t += 1                                                      # tradam.py:82
        loss, grad = fg(w)
        grad = grad.to(self.m.dtype)
        self.m.mul_(self.betas[0]).add_(grad, alpha=1 - self.betas[0])   # :87
        self.v.mul_(self.betas[1]).addcmul_(grad, grad, value=1 - self.betas[1])   # :88
        m_hat = self.m / (1 - self.betas[0] ** self.t)                   # :91
        v_hat = self.v / (1 - self.betas[1] ** self.t)                   # :92
        s = m_hat.div(v_hat.sqrt().add(self.eps))                        # :95
        if self.norm_type == torch.inf:                                  # :98-101
            step_len = torch.norm(s, p=torch.inf).item()
        else:
            step_len = s.dot(s).item()                                   # (sic) squared 2-norm
        self.step_len = step_len
        if step_len > self.lr:                                           # :104-107
            self.step_buf.copy_(s.mul(self.lr / step_len))
        else:
            self.step_buf.copy_(s.mul(self.lr))
        w.add_(self.step_buf.to(w.dtype) * -1.0)                         # :110, :77-78
        return loss
