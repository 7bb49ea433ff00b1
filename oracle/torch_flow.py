"""Stock-PyTorch restatement of the reference coupling flow (eager ops + autograd), device-agnostic.

TEST INFRASTRUCTURE ONLY, like flow_oracle.py: it is the "what stock PyTorch does on the same GPU" bar
(tools/eager_bar.py) and a second checker for the GPU tests at sizes where the numpy oracle is slow.
Nothing in the shipped package imports it.

It composes the same torch ops the reference module is made of (nn.Linear / LeakyReLU / LayerNorm, a dense
x @ W, triangular solves for the inverse), takes the reference's state_dict keys, and is pinned by
tests/test_oracle_golden.py against the fixtures generated from the real reference module.

Reference: /root/reference/il-and-cil/utils/torch_fcnf.py:12-152 (flax deltas nf.py:94-104,131-148 behind
the first_hidden / ln_eps arguments).
"""
from __future__ import annotations

import torch
from torch import nn


def _conditioner(k_in, h1, c, out, eps):
    # torch_fcnf.py:73-79: Linear, LeakyReLU, LayerNorm, Linear, LeakyReLU, LayerNorm, Linear
    return nn.Sequential(nn.Linear(k_in, h1), nn.LeakyReLU(), nn.LayerNorm(h1, eps=eps),
                         nn.Linear(h1, c), nn.LeakyReLU(), nn.LayerNorm(c, eps=eps), nn.Linear(c, out))


class _PLU(nn.Module):
    def __init__(self, D):
        super().__init__()
        for name in ("U", "L"):
            setattr(self, name, nn.Parameter(torch.zeros(D, D)))
        self.s = nn.Parameter(torch.ones(D))
        for name in ("P", "P_inv"):
            setattr(self, name, nn.Parameter(torch.eye(D), requires_grad=False))

    def factors(self):
        D = self.s.numel()
        eye = torch.eye(D, dtype=self.s.dtype, device=self.s.device)
        return self.L.tril(-1) + eye, self.U.triu(1) + torch.diag(self.s), eye

    def weight(self):            # torch_fcnf.py:34-38
        Lh, Uh, _ = self.factors()
        return self.P @ Lh @ Uh

    def weight_inv(self):        # torch_fcnf.py:45-54
        Lh, Uh, eye = self.factors()
        Ui = torch.linalg.solve_triangular(Uh, eye, upper=True)
        Li = torch.linalg.solve_triangular(Lh, eye, upper=False, unitriangular=True)
        return Ui @ Li @ self.P_inv

    def logdet(self):            # torch_fcnf.py:41
        return self.s.abs().log().sum()


class _Block(nn.Module):
    def __init__(self, D, C, R, H1, eps):
        super().__init__()
        self.dc = (D + 1) // 2
        self.l = _PLU(D)
        self.t = _conditioner(self.dc + R, H1, C, D // 2, eps)
        self.s = _conditioner(self.dc + R, H1, C, D // 2, eps)

    def forward(self, x, y):     # torch_fcnf.py:95-106
        x = x @ self.l.weight()
        xc, xt = x[:, :self.dc], x[:, self.dc:]
        h = torch.cat([xc, y], 1)
        s, t = self.s(h), self.t(h)
        return torch.cat([xc, (xt - t) * torch.exp(-s)], 1), self.l.logdet() - s.sum(1)

    def reverse(self, z, y, with_logdet=False):     # torch_fcnf.py:108-117; log-det: flax nf.py:131-148
        zc, zt = z[:, :self.dc], z[:, self.dc:]
        h = torch.cat([zc, y], 1)
        s, t = self.s(h), self.t(h)
        x = torch.cat([zc, zt * torch.exp(s) + t], 1) @ self.l.weight_inv()
        return (x, s.sum(1) - self.l.logdet()) if with_logdet else x


class TorchFlow(nn.Module):
    """forward(x, y) -> (z, outputs, log_dets); reverse(z, y) -> x   (torch_fcnf.py:119-152)."""

    def __init__(self, D, C, R, n, first_hidden=None, ln_eps=1e-5):
        super().__init__()
        self.blocks = nn.ModuleList(_Block(D, C, R, first_hidden or C, ln_eps) for _ in range(n))

    def forward(self, x, y):
        logdets = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
        outs = []
        for b in self.blocks:
            x, ld = b(x, y)
            logdets = logdets + ld
            outs.append(x)
        return x, outs, logdets

    def reverse(self, z, y, return_logdet=False):
        logdets = torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
        for b in reversed(self.blocks):
            if return_logdet:
                z, ld = b.reverse(z, y, True)
                logdets = logdets + ld
            else:
                z = b.reverse(z, y)
        return (z, logdets) if return_logdet else z


class TorchEncoder(nn.Module):
    """Stock-PyTorch restatement of the reference conditioning encoder (torch_fcnf.py:154-190): n_hidden times
    Linear -> LayerNorm -> SiLU, then Linear; same state_dict keys (fc1.., norm1.., fc_out)."""

    def __init__(self, input_size, rep_size, hidden=512, n_hidden=4, ln_eps=1e-5):
        super().__init__()
        self.n_hidden = n_hidden
        fan = input_size
        for i in range(1, n_hidden + 1):
            setattr(self, f"fc{i}", nn.Linear(fan, hidden))
            fan = hidden
        self.fc_out = nn.Linear(hidden, rep_size)
        for i in range(1, n_hidden + 1):
            setattr(self, f"norm{i}", nn.LayerNorm(hidden, eps=ln_eps))

    def forward(self, g):
        x = g
        for i in range(1, self.n_hidden + 1):
            x = nn.functional.silu(getattr(self, f"norm{i}")(getattr(self, f"fc{i}")(x)))
        return self.fc_out(x)
