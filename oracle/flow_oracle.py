"""CPU restatement (numpy) of the reference coupling flow -- the parity ORACLE.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package imports this module; it is
used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl
reference legs, always as the checker or the CPU baseline, never as the product.

Parity pin: this restatement is checked against the real reference module
(/root/reference/il-and-cil/utils/torch_fcnf.py, imported in the build container by
tests/golden/make_golden.py) through the fixtures committed under tests/golden/
(tests/test_oracle_golden.py).  The flax-variant flags (first hidden width C+R, LayerNorm
eps 1e-6, reverse log-det; /root/reference/unsupservised-gcrl/nf.py:94-104,131-148) are
"parity unpinned": jax/flax are not installable here, so they are pinned only
transitively (with H1=C, eps=1e-5 they reduce to the torch path).

Every function cites the reference lines it follows.  Weights are a dict with the
reference's state_dict keys (numpy arrays, Linear weight (out, in)).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import scipy.linalg as sla

LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default, torch_fcnf.py:75


# --------------------------------------------------------------------------- PLU
def plu_weight(sd, i):
    """W = P @ (tril(L,-1)+I) @ (triu(U,1)+diag(s))  -- torch_fcnf.py:34-38."""
    p = f"blocks.{i}.l."
    D = sd[p + "s"].shape[0]
    Lh = np.tril(sd[p + "L"], -1) + np.eye(D, dtype=sd[p + "L"].dtype)
    Uh = np.triu(sd[p + "U"], 1) + np.diag(sd[p + "s"])
    return sd[p + "P"] @ Lh @ Uh, Lh, Uh


def plu_weight_inv(sd, i):
    """W^-1 = U^-1 @ L^-1 @ P_inv with both inverses obtained by triangular solves
    against the identity -- torch_fcnf.py:45-54 (algorithm kept, not a nicer one)."""
    p = f"blocks.{i}.l."
    _, Lh, Uh = plu_weight(sd, i)
    D = Lh.shape[0]
    eye = np.eye(D, dtype=Lh.dtype)
    U_inv = sla.solve_triangular(Uh, eye, lower=False)
    L_inv = sla.solve_triangular(Lh, eye, lower=True, unit_diagonal=True)
    return (U_inv @ L_inv @ sd[p + "P_inv"]).astype(Lh.dtype)


def plu_logdet(sd, i):
    """sum(log|s|)  -- torch_fcnf.py:41."""
    s = sd[f"blocks.{i}.l.s"]
    return np.sum(np.log(np.abs(s)))


# --------------------------------------------------------------------------- MLP pieces
def leaky_relu(u):
    return np.where(u > 0, u, LEAKY_SLOPE * u)


# flax nn.LayerNorm (unsupservised-gcrl/nf.py:94-104; flax 0.8.3 default use_fast_variance=True) computes the variance
# as max(0, E[x^2] - E[x]^2); torch nn.LayerNorm (torch_fcnf.py:75) the centred form.  The two are the same function
# (and have the same derivative); they differ in rounding only.  Tests switch this flag to check the product's
# `ln_fast_variance` option against the matching formula (parity unpinned: no jax in this image).
LN_FAST_VARIANCE = False


def layer_norm_stats(v, eps):
    """mean and 1/sqrt(var + eps) over the last axis (biased variance, as both frameworks use)."""
    mu = v.mean(axis=1, keepdims=True)
    if LN_FAST_VARIANCE:
        var = np.maximum((v * v).mean(axis=1, keepdims=True) - mu * mu, 0)
    else:
        var = ((v - mu) ** 2).mean(axis=1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + v.dtype.type(eps))
    return mu, rstd


def conditioner_forward(sd, prefix, h0, eps, keep=False):
    """Linear-LeakyReLU-LayerNorm x2 then Linear -- torch_fcnf.py:74-78 (t), :85-89 (s)."""
    W1, b1 = sd[prefix + "0.weight"], sd[prefix + "0.bias"]
    g1, be1 = sd[prefix + "2.weight"], sd[prefix + "2.bias"]
    W2, b2 = sd[prefix + "3.weight"], sd[prefix + "3.bias"]
    g2, be2 = sd[prefix + "5.weight"], sd[prefix + "5.bias"]
    W3, b3 = sd[prefix + "6.weight"], sd[prefix + "6.bias"]
    u1 = h0 @ W1.T + b1
    v1 = leaky_relu(u1)
    mu1, r1 = layer_norm_stats(v1, eps)
    xh1 = (v1 - mu1) * r1
    h1 = xh1 * g1 + be1
    u2 = h1 @ W2.T + b2
    v2 = leaky_relu(u2)
    mu2, r2 = layer_norm_stats(v2, eps)
    xh2 = (v2 - mu2) * r2
    h2 = xh2 * g2 + be2
    out = h2 @ W3.T + b3
    if keep:
        return out, dict(h0=h0, u1=u1, xh1=xh1, r1=r1, h1=h1, u2=u2, xh2=xh2, r2=r2, h2=h2)
    return out


def _ln_lrelu_backward(dh, gamma, xh, rstd, u):
    """Backward of LayerNorm(LeakyReLU(u)) given d(output)."""
    dxh = dh * gamma
    m1 = dxh.mean(axis=1, keepdims=True)
    m2 = (dxh * xh).mean(axis=1, keepdims=True)
    dv = rstd * (dxh - m1 - xh * m2)
    return dv * np.where(u > 0, 1.0, LEAKY_SLOPE).astype(dv.dtype)


def conditioner_backward(sd, prefix, cache, dout, grads):
    """Hand-derived backward of conditioner_forward (checked against torch.autograd of
    the reference through the golden fixtures).  Returns d(h0)."""
    W1 = sd[prefix + "0.weight"]
    g1 = sd[prefix + "2.weight"]
    W2 = sd[prefix + "3.weight"]
    g2 = sd[prefix + "5.weight"]
    W3 = sd[prefix + "6.weight"]
    c = cache
    grads[prefix + "6.weight"] = dout.T @ c["h2"]
    grads[prefix + "6.bias"] = dout.sum(0)
    dh2 = dout @ W3
    grads[prefix + "5.weight"] = (dh2 * c["xh2"]).sum(0)
    grads[prefix + "5.bias"] = dh2.sum(0)
    du2 = _ln_lrelu_backward(dh2, g2, c["xh2"], c["r2"], c["u2"])
    grads[prefix + "3.weight"] = du2.T @ c["h1"]
    grads[prefix + "3.bias"] = du2.sum(0)
    dh1 = du2 @ W2
    grads[prefix + "2.weight"] = (dh1 * c["xh1"]).sum(0)
    grads[prefix + "2.bias"] = dh1.sum(0)
    du1 = _ln_lrelu_backward(dh1, g1, c["xh1"], c["r1"], c["u1"])
    grads[prefix + "0.weight"] = du1.T @ c["h0"]
    grads[prefix + "0.bias"] = du1.sum(0)
    return du1 @ W1


# --------------------------------------------------------------------------- blocks
def block_forward(sd, i, x, y, eps=1e-5, keep=False):
    """MetaBlock.forward -- torch_fcnf.py:95-106.  Returns (x_out, logdet (B,))."""
    D = x.shape[1]
    dc = (D + 1) // 2  # torch.tensor_split(x, 2, dim=1): first chunk gets the extra column
    W, _, _ = plu_weight(sd, i)
    xp = x @ W                                             # :40
    ld0 = plu_logdet(sd, i)                                # :41
    xc, xt = xp[:, :dc], xp[:, dc:]                        # :98
    h0 = np.concatenate([xc, y], axis=1)                   # :99-100
    p = f"blocks.{i}."
    if keep:
        s, cs = conditioner_forward(sd, p + "s.", h0, eps, keep=True)
        t, ct = conditioner_forward(sd, p + "t.", h0, eps, keep=True)
    else:
        s = conditioner_forward(sd, p + "s.", h0, eps)
        t = conditioner_forward(sd, p + "t.", h0, eps)
    zt = (xt - t) * np.exp(-s)                             # :101
    z = np.concatenate([xc, zt], axis=1)                   # :102
    ld = ld0 - s.sum(axis=1)                               # :104
    if keep:
        return z, ld.astype(x.dtype), dict(x=x, W=W, s=s, t=t, zt=zt, cs=cs, ct=ct)
    return z, ld.astype(x.dtype)


def block_reverse(sd, i, z, y, eps=1e-5):
    """MetaBlock.reverse -- torch_fcnf.py:108-117 (affine first, PLU^-1 second).
    Also returns the flax-style reverse log-det  -sum log|s_plu| + sum_j s_j
    (unsupservised-gcrl/nf.py:131-148)."""
    D = z.shape[1]
    dc = (D + 1) // 2
    zc, zt = z[:, :dc], z[:, dc:]
    h0 = np.concatenate([zc, y], axis=1)
    p = f"blocks.{i}."
    s = conditioner_forward(sd, p + "s.", h0, eps)
    t = conditioner_forward(sd, p + "t.", h0, eps)
    xt = zt * np.exp(s) + t                                # :112
    xp = np.concatenate([zc, xt], axis=1)                  # :113
    x = xp @ plu_weight_inv(sd, i)                         # :115 -> :44-58
    ld = -plu_logdet(sd, i) + s.sum(axis=1)
    return x.astype(z.dtype), ld.astype(z.dtype)


def flow_forward(sd, n, x, y, eps=1e-5, keep=False):
    """RealNVP.forward -- torch_fcnf.py:134-141.  Returns (z, outputs list, log_dets)."""
    outputs, caches = [], []
    log_dets = np.zeros(x.shape[0], dtype=x.dtype)
    for i in range(n):
        if keep:
            x, ld, c = block_forward(sd, i, x, y, eps, keep=True)
            caches.append(c)
        else:
            x, ld = block_forward(sd, i, x, y, eps)
        log_dets = log_dets + ld
        outputs.append(x)
    if keep:
        return x, outputs, log_dets, caches
    return x, outputs, log_dets


def flow_reverse(sd, n, z, y, eps=1e-5, return_sequence=False, return_logdet=False):
    """RealNVP.reverse -- torch_fcnf.py:143-152."""
    seq = [z]
    lds = np.zeros(z.shape[0], dtype=z.dtype)
    for i in reversed(range(n)):
        z, ld = block_reverse(sd, i, z, y, eps)
        lds = lds + ld
        seq.append(z)
    out = seq if return_sequence else z
    if return_logdet:
        return out, lds
    return out


def std_normal_log_prob(z):
    """MultivariateNormal(0, I).log_prob(z) (torch_nfbc_ant.py:146,269)."""
    D = z.shape[1]
    return -0.5 * (z * z).sum(axis=1) - 0.5 * D * np.log(2.0 * np.pi)


def flow_backward(sd, n, x, y, dz, dld, eps=1e-5):
    """Gradients of  sum(z*dz) + sum(log_dets*dld)  w.r.t. x, y and every trainable
    parameter (what loss.backward() produces for torch_nfbc_ant.py:269-272 when
    dz = d loss/dz and dld = d loss/d log_dets)."""
    z, outputs, log_dets, caches = flow_forward(sd, n, x, y, eps, keep=True)
    D = x.shape[1]
    dc = (D + 1) // 2
    grads: "OrderedDict[str, np.ndarray]" = OrderedDict()
    dy = np.zeros_like(y)
    dld_col = dld[:, None]
    for i in reversed(range(n)):
        c = caches[i]
        p = f"blocks.{i}."
        dzc, dzt = dz[:, :dc], dz[:, dc:]
        e = np.exp(-c["s"])
        dxt = dzt * e
        dt_ = -dzt * e
        ds_ = -dzt * c["zt"] - dld_col
        dh0 = conditioner_backward(sd, p + "s.", c["cs"], ds_, grads)
        dh0 = dh0 + conditioner_backward(sd, p + "t.", c["ct"], dt_, grads)
        dy += dh0[:, dc:]
        dxp = np.concatenate([dzc + dh0[:, :dc], dxt], axis=1)
        # invertible linear layer: xp = x @ W, W = P Lh Uh
        W, Lh, Uh = plu_weight(sd, i)
        dW = c["x"].T @ dxp
        P = sd[p + "l.P"]
        dLh = P.T @ dW @ Uh.T
        dUh = (P @ Lh).T @ dW
        grads[p + "l.L"] = np.tril(dLh, -1)
        grads[p + "l.U"] = np.triu(dUh, 1)
        grads[p + "l.s"] = np.diag(dUh) + dld.sum() / sd[p + "l.s"]
        dz = dxp @ W.T
    return dict(dx=dz, dy=dy, grads=grads, z=z, log_dets=log_dets, outputs=outputs)


def cast_state_dict(sd, dtype):
    return OrderedDict((k, np.asarray(v, dtype=dtype)) for k, v in sd.items())


def min_abs_preactivation(sd, n, x, y, eps=1e-5):
    """Smallest |u| over every LeakyReLU input of the stack (float64 recommended).  LeakyReLU is
    not differentiable at 0: when some |u| is within float32 rounding of zero, two correct float32
    implementations can pick different slopes and their gradients differ by O(1e-3).  Parity cases
    are therefore drawn so that this margin is comfortably above float32 noise."""
    _, _, _, caches = flow_forward(sd, n, x, y, eps, keep=True)
    m = np.inf
    for c in caches:
        for key in ("cs", "ct"):
            m = min(m, float(np.abs(c[key]["u1"]).min()), float(np.abs(c[key]["u2"]).min()))
    return m
