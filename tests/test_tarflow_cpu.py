"""Causal-transformer flow, CPU part: the checker's own invariants (parity is unpinned
upstream, so the restatement is validated by inverse-of-forward and by the Jacobian log-determinant), and the host-side
mirror's contract with it (state_dict keys, same-seed initialisation, the reference's constructor and import path)."""
import numpy as np
import pytest
import torch

import nf_b200
from oracle import torch_tarflow as tt

CASES = [dict(A=8, C=64, R=16, hd=32, nb=3, L=2, clamp=0.0), dict(A=5, C=32, R=7, hd=32, nb=2, L=1, clamp=1.5),
         dict(A=1, C=32, R=4, hd=32, nb=2, L=1, clamp=0.0)]


def make_oracle(c, seed=0, dtype=torch.float64):
    torch.manual_seed(seed)
    m = tt.RealNVP(c["A"], c["C"], c["R"], c["hd"], c["nb"], c["L"], None, logscale_clamp=c["clamp"])
    tt.perturb_output_heads(m, seed=seed + 1)
    return m.to(dtype)


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"A{c['A']}C{c['C']}")
def test_checker_inverse_and_jacobian_logdet(c):
    m = make_oracle(c)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, c["A"], 1, generator=g, dtype=torch.float64)
    y = torch.randn(6, 1, c["R"], generator=g, dtype=torch.float64)
    z, outs, ld = m(x, y)
    assert len(outs) == c["nb"] and z.shape == x.shape and ld.shape == (6,)
    assert torch.equal(outs[-1], z)
    assert float((m.reverse(z, y) - x).abs().max()) < 1e-10
    seq = m.reverse(z, y, return_sequence=True)
    assert len(seq) == c["nb"] + 1 and float((seq[1] - outs[-2]).abs().max()) < 1e-10 if c["nb"] > 1 else True
    for b in range(2):       # log|det dz/dx| from autograd equals the reported log-det
        J = torch.autograd.functional.jacobian(lambda v: m(v.reshape(1, c["A"], 1), y[b:b + 1])[0].reshape(-1), x[b].reshape(-1))
        assert abs(float(torch.slogdet(J)[1]) - float(ld[b])) < 1e-9
    # autoregressive structure inside a block: output dimension t depends on inputs <= t only (in the block's order)
    blk = m.blocks[0]
    J = torch.autograd.functional.jacobian(lambda v: blk(v.reshape(1, c["A"], 1), y[:1])[0].reshape(-1), x[0].reshape(-1))
    assert float(torch.triu(J, diagonal=1).abs().max()) == 0.0


def test_fresh_model_is_the_identity():
    torch.manual_seed(0)
    m = tt.RealNVP(8, 64, 16, 32, 2, 1, None)
    x, y = torch.randn(4, 8, 1), torch.randn(4, 1, 16)
    z, _, ld = m(x, y)
    assert torch.equal(z, x) and float(ld.abs().max()) == 0.0


def test_mirror_has_the_checkers_keys_and_initialisation():
    for c in CASES[:2]:
        torch.manual_seed(3)
        ours = nf_b200.TransformerRealNVP(c["A"], c["C"], c["R"], c["hd"], c["nb"], c["L"], None)
        torch.manual_seed(3)
        ref = tt.RealNVP(c["A"], c["C"], c["R"], c["hd"], c["nb"], c["L"], None)
        sa, sb = ours.state_dict(), ref.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        assert all(sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]) for k in sa)
        # the parameters are views of one flat arena, and loading a checkpoint keeps them there
        ours.load_state_dict({k: v + 1.0 for k, v in sb.items()})
        assert ours._is_flat()
        assert torch.equal(ours.blocks[0].layers[0].attn.qkv.weight, sb["blocks.0.layers.0.attn.qkv.weight"] + 1.0)
        n_params = sum(p.numel() for p in ref.parameters())
        assert n_params <= ours._total < n_params + 4 * len(sa)        # slots are padded to 16 bytes


def test_reference_import_path_and_argument_errors():
    from nf_b200.torch_networks import GEncoder, RealNVP        # torch_nfgcbc.py:27
    assert RealNVP is nf_b200.TransformerRealNVP and GEncoder is nf_b200.GEncoder
    m = RealNVP(8, 256, 64, 64, 4, 4, None)                     # torch_nfgcbc.py:119 with the script's defaults (:77-82)
    assert len(m.blocks) == 4 and len(m.blocks[0].layers) == 4
    with pytest.raises(RuntimeError):                           # no CPU fallback
        m(torch.zeros(2, 8, 1), torch.zeros(2, 1, 64))
    with pytest.raises(nf_b200._lib.NfError):
        RealNVP(8, 250, 64, 64, 4, 4, None)                     # head_dim must divide channels
    with pytest.raises(nf_b200._lib.NfError):
        RealNVP(64, 256, 64, 64, 4, 4, None)                    # at most 32 dimensions (lane = key in the softmax)
