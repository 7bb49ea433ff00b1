"""TEST INFRASTRUCTURE -- CPU/GPU checker for the causal-transformer flow.

PARITY UNPINNED.  The reference imports this model from `utils.torch_networks` (`il-and-cil/torch_nfgcbc.py:27`), a file
that is absent from the reference tree; upstream says it follows apple/ml-tarflow (`README.md:26`), which is not vendored
and not pinned.  What the reference does pin is the call-site contract, and this restatement follows it:

  * constructor  RealNVP(input_dims, channels, rep_dim, head_dim, blocks, layers_per_block, prior)   torch_nfgcbc.py:119
  * forward      model(actions (B, A, 1), cond (B, 1, rep)) -> (z (B, A, 1), outputs [blocks x (B, A, 1)], logdets (B,))
                 :136-142, loss -(prior.log_prob(z.squeeze(-1)) + logdets).mean() :142, outputs[b].pow(2).mean() :155
  * reverse      model.reverse(prior_sample (B, A, 1), cond) -> actions (B, A, 1)          utils/torch_evaluations.py:79-82

The block is the published TarFlow meta block (Zhai et al., "Normalizing Flows are Capable Generative Models"), restated
for a continuous conditioning vector: tokens = ONE conditioning token (`cond_proj(cond)`, which is why the caller passes
cond as (B, 1, rep)) followed by the first A - 1 action dimensions (1 channel each; the last dimension conditions nothing,
so it is never embedded: T = A tokens), learned positional embedding, `layers_per_block`
pre-LayerNorm blocks [causal multi-head attention (head_dim) + GELU MLP (expansion 4)], a zero-initialised output
projection to (log-scale xa, shift xb) per token.  Output position t (the conditioning token for t = 0, action token
t - 1 otherwise) parameterises action dimension t, so the Jacobian is triangular:
    z_t = (x_t - xb_t) exp(-xa_t),   logdet = -sum_t xa_t,   inverse: x_t = z_t exp(xa_t) + xb_t for t = 0 .. A-1 in order.
Odd blocks see the action dimensions in reversed order (TarFlow's flip permutation) and hand z back un-flipped.
`logscale_clamp = c > 0` bounds the log-scale as xa = c tanh(raw / c) (BASELINE.json north_star item 2); 0 = raw.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import math

import torch
from torch import nn
import torch.nn.functional as F


class Attention(nn.Module):
    def __init__(self, channels, head_dim):
        super().__init__()
        assert channels % head_dim == 0
        self.norm = nn.LayerNorm(channels)
        self.qkv = nn.Linear(channels, 3 * channels)
        self.proj = nn.Linear(channels, channels)
        self.heads, self.head_dim = channels // head_dim, head_dim

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = self.qkv(self.norm(x)).reshape(B, T, 3, self.heads, self.head_dim).permute(2, 0, 3, 1, 4)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(self.head_dim)
        mask = torch.ones(T, T, dtype=torch.bool, device=x.device).tril()
        p = s.masked_fill(~mask, float("-inf")).softmax(-1)
        o = (p @ v).transpose(1, 2).reshape(B, T, C)
        return self.proj(o)


class MLP(nn.Module):
    def __init__(self, channels, expansion):
        super().__init__()
        self.norm = nn.LayerNorm(channels)
        self.fc1 = nn.Linear(channels, channels * expansion)
        self.fc2 = nn.Linear(channels * expansion, channels)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(self.norm(x))))


class AttentionBlock(nn.Module):
    def __init__(self, channels, head_dim, expansion):
        super().__init__()
        self.attn = Attention(channels, head_dim)
        self.mlp = MLP(channels, expansion)

    def forward(self, x):
        x = x + self.attn(x)
        return x + self.mlp(x)


class MetaBlock(nn.Module):
    def __init__(self, A, channels, rep_dim, head_dim, layers, flip, expansion=4, logscale_clamp=0.0):
        super().__init__()
        self.A, self.flip, self.clamp = A, flip, float(logscale_clamp)
        self.proj_in = nn.Linear(1, channels)
        self.cond_proj = nn.Linear(rep_dim, channels)
        self.pos_embed = nn.Parameter(torch.randn(A, channels) * 1e-2)
        self.layers = nn.ModuleList([AttentionBlock(channels, head_dim, expansion) for _ in range(layers)])
        self.proj_out = nn.Linear(channels, 2)
        nn.init.zeros_(self.proj_out.weight)
        nn.init.zeros_(self.proj_out.bias)

    def params_of(self, xp, cond):
        """(xa, xb), each (B, A, 1): output position t sees the conditioning token and xp[:, :t] only."""
        tok = torch.cat([self.cond_proj(cond), self.proj_in(xp[:, : self.A - 1])], dim=1) + self.pos_embed
        for layer in self.layers:
            tok = layer(tok)
        out = self.proj_out(tok)
        xa, xb = out[..., 0:1], out[..., 1:2]
        if self.clamp > 0:
            xa = self.clamp * torch.tanh(xa / self.clamp)
        return xa, xb

    def forward(self, x, cond):
        xp = x.flip(1) if self.flip else x
        xa, xb = self.params_of(xp, cond)
        zp = (xp - xb) * torch.exp(-xa)
        return (zp.flip(1) if self.flip else zp), -xa.sum(dim=(1, 2))

    def reverse(self, z, cond):
        zp = z.flip(1) if self.flip else z
        xp = torch.zeros_like(zp)
        for t in range(self.A):      # dimension t needs x_0 .. x_{t-1}: the autoregressive loop
            xa, xb = self.params_of(xp, cond)
            xp = torch.cat([xp[:, :t], zp[:, t:t + 1] * torch.exp(xa[:, t:t + 1]) + xb[:, t:t + 1], xp[:, t + 1:]], dim=1)
        return xp.flip(1) if self.flip else xp


class RealNVP(nn.Module):
    """Constructor order of torch_nfgcbc.py:119."""

    def __init__(self, in_channels, channels, rep_dim, head_dim, blocks, layers_per_block, prior=None, expansion=4,
                 logscale_clamp=0.0):
        super().__init__()
        self.prior = prior
        self.blocks = nn.ModuleList([MetaBlock(in_channels, channels, rep_dim, head_dim, layers_per_block, flip=(i % 2 == 1),
                                               expansion=expansion, logscale_clamp=logscale_clamp) for i in range(blocks)])

    def forward(self, x, cond):
        log_dets = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
        outputs = []
        for block in self.blocks:
            x, ld = block(x, cond)
            log_dets = log_dets + ld
            outputs.append(x)
        return x, outputs, log_dets

    def reverse(self, z, cond, return_sequence=False):
        seq = [z]
        for block in reversed(self.blocks):
            z = block.reverse(z, cond)
            seq.append(z)
        return seq if return_sequence else z


def perturb_output_heads(model, std=None, seed=0):
    """The output projections are zero-initialised (a fresh model is the identity): parity inputs perturb them."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for b in model.blocks:
            C = b.proj_out.weight.shape[1]
            s = std if std is not None else 0.5 / math.sqrt(C)
            b.proj_out.weight.copy_((torch.randn(b.proj_out.weight.shape, generator=g) * s).to(b.proj_out.weight))
            b.proj_out.bias.copy_((torch.randn(2, generator=g) * 0.05).to(b.proj_out.bias))
