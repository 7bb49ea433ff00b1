"""CPU restatement (numpy) of the reference conditioning encoder -- parity oracle for SURVEY.md section 8 row f2.

TEST INFRASTRUCTURE ONLY (see flow_oracle.py).  Pinned against the real reference module
(/root/reference/il-and-cil/utils/torch_fcnf.py:154-190, imported by tests/golden/make_golden_encoder.py) through
the fixtures under tests/golden/encoder/.  The flax twin (nf.py:188-212) differs by LayerNorm eps = 1e-6: that
flag is "parity unpinned" like the flow's flax flags.

Weights are a dict with the reference's state_dict keys: fc1..fcN.{weight,bias}, norm1..normN.{weight,bias},
fc_out.{weight,bias}; Linear weight is (out, in).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def n_hidden_layers(sd):
    n = 0
    while f"fc{n + 1}.weight" in sd:
        n += 1
    return n


def sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


def encoder_forward(sd, g, eps=1e-5, keep=False):
    """x = fc_i(x); x = norm_i(x); x = silu(x) for every hidden layer, then fc_out  (torch_fcnf.py:170-188)."""
    h = g
    cache = []
    for i in range(1, n_hidden_layers(sd) + 1):
        u = h @ sd[f"fc{i}.weight"].T + sd[f"fc{i}.bias"]
        mean = u.mean(axis=1, keepdims=True)
        var = ((u - mean) ** 2).mean(axis=1, keepdims=True)            # biased, like nn.LayerNorm
        rstd = 1.0 / np.sqrt(var + u.dtype.type(eps))
        xh = (u - mean) * rstd
        a = xh * sd[f"norm{i}.weight"] + sd[f"norm{i}.bias"]
        out = a * sigmoid(a)                                            # F.silu
        cache.append((h, xh, rstd, a))
        h = out
    y = h @ sd["fc_out.weight"].T + sd["fc_out.bias"]
    return (y, cache, h) if keep else y


def encoder_backward(sd, g, dout, eps=1e-5):
    """Gradients of sum(out * dout) w.r.t. the input and every parameter."""
    y, cache, h_last = encoder_forward(sd, g, eps, keep=True)
    grads = OrderedDict()
    grads["fc_out.weight"] = dout.T @ h_last
    grads["fc_out.bias"] = dout.sum(axis=0)
    dh = dout @ sd["fc_out.weight"]
    for i in range(n_hidden_layers(sd), 0, -1):
        h_in, xh, rstd, a = cache[i - 1]
        sg = sigmoid(a)
        da = dh * sg * (1.0 + a * (1.0 - sg))
        grads[f"norm{i}.weight"] = (da * xh).sum(axis=0)
        grads[f"norm{i}.bias"] = da.sum(axis=0)
        dxh = da * sd[f"norm{i}.weight"]
        du = rstd * (dxh - dxh.mean(axis=1, keepdims=True) - xh * (dxh * xh).mean(axis=1, keepdims=True))
        grads[f"fc{i}.weight"] = du.T @ h_in
        grads[f"fc{i}.bias"] = du.sum(axis=0)
        dh = du @ sd[f"fc{i}.weight"]
    return dict(out=y, dg=dh, grads=grads)
