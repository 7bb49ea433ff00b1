"""Deterministic synthetic weights and inputs for the coupling-flow hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything here is generated with
numpy's legacy ``RandomState`` and elementwise IEEE arithmetic only (no BLAS, no
LAPACK, no torch RNG), so the same (config, seed) gives bit-identical arrays in
this container, on the GPU box and inside the script that produced the golden
fixtures under tests/golden/.

The tensors use the reference's ``state_dict`` key layout
(/root/reference/il-and-cil/utils/torch_fcnf.py:14-31 for ``blocks.i.l.*``,
:74-93 for ``blocks.i.{t,s}.{0,2,3,5,6}.*``).  The reference zero-initialises the
last Linear of both conditioners (:79-82, :90-93); here they are drawn from
N(0, (0.2/sqrt(C))^2) so that parity checks are not vacuous.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np


def dims(D: int, C: int, R: int, H1: int | None = None):
    """Split sizes used everywhere: torch.tensor_split(x, 2, dim=1)
    (torch_fcnf.py:98) gives ceil(D/2) conditioning and D//2 transformed dims."""
    dc = (D + 1) // 2
    dt = D // 2
    H1 = C if H1 is None else H1
    return dc, dt, H1


def make_state_dict(D: int, C: int, R: int, n: int, seed: int = 0, H1: int | None = None,
                    last_std_scale: float = 0.2, dtype=np.float32) -> "OrderedDict[str, np.ndarray]":
    """Weights with the reference's key names and shapes (Linear weight is (out, in))."""
    dc, dt, H1 = dims(D, C, R, H1)
    K1 = dc + R
    rs = np.random.RandomState(seed)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def uni(shape, fan_in):
        k = 1.0 / math.sqrt(fan_in)
        return rs.uniform(-k, k, size=shape)

    for i in range(n):
        p = f"blocks.{i}."
        # PLU factors: P a permutation matrix, L unit-lower, U strictly-upper, s the diagonal.
        perm = rs.permutation(D)
        P = np.zeros((D, D))
        P[np.arange(D), perm] = 1.0
        amp = 0.6 / math.sqrt(D)
        L = np.tril(rs.standard_normal((D, D)) * amp, -1) + np.eye(D)
        U = np.triu(rs.standard_normal((D, D)) * amp, 1)
        s = rs.uniform(0.6, 1.4, size=D) * np.where(rs.uniform(size=D) < 0.5, -1.0, 1.0)
        sd[p + "l.s"] = s
        # The reference stores full matrices and masks at use (torch_fcnf.py:34-35); put junk in the
        # masked-out triangles so that the masking itself is exercised.
        sd[p + "l.U"] = U + np.tril(rs.standard_normal((D, D)) * 0.05, 0)
        sd[p + "l.L"] = L + np.triu(rs.standard_normal((D, D)) * 0.05, 1)
        sd[p + "l.P"] = P
        sd[p + "l.P_inv"] = P.T.copy()
        for net in ("t", "s"):
            q = p + net + "."
            sd[q + "0.weight"] = uni((H1, K1), K1)
            sd[q + "0.bias"] = uni((H1,), K1)
            sd[q + "2.weight"] = 1.0 + 0.1 * rs.standard_normal(H1)
            sd[q + "2.bias"] = 0.1 * rs.standard_normal(H1)
            sd[q + "3.weight"] = uni((C, H1), H1)
            sd[q + "3.bias"] = uni((C,), H1)
            sd[q + "5.weight"] = 1.0 + 0.1 * rs.standard_normal(C)
            sd[q + "5.bias"] = 0.1 * rs.standard_normal(C)
            sd[q + "6.weight"] = rs.standard_normal((dt, C)) * (last_std_scale / math.sqrt(C))
            sd[q + "6.bias"] = rs.standard_normal((dt,)) * (last_std_scale / math.sqrt(C))
    return OrderedDict((k, np.ascontiguousarray(v, dtype=dtype)) for k, v in sd.items())


def make_inputs(B: int, D: int, R: int, seed: int = 1234, dtype=np.float32):
    """x = U(-1,1) + 0.1 N(0,1) (normalised actions plus the training noise of
    torch_nfbc_ant.py:133-138,264-265); y = N(0,1) (encoder output stand-in)."""
    rs = np.random.RandomState(seed)
    x = rs.uniform(-1.0, 1.0, size=(B, D)) + 0.1 * rs.standard_normal((B, D))
    y = rs.standard_normal((B, R))
    return np.ascontiguousarray(x, dtype=dtype), np.ascontiguousarray(y, dtype=dtype)


def make_cotangents(B: int, D: int, seed: int = 4321, dtype=np.float32):
    """Random upstream gradients (dz, dlogdet) for backward parity checks."""
    rs = np.random.RandomState(seed)
    dz = rs.standard_normal((B, D))
    dld = rs.standard_normal((B,))
    return np.ascontiguousarray(dz, dtype=dtype), np.ascontiguousarray(dld, dtype=dtype)


# BASELINE.json configs as (D, C, R, n, H1, ln_eps) -- SURVEY.md section 8(d).
CONFIGS = {
    "C1a": dict(D=8, C=512, R=256, n=12, H1=512, ln_eps=1e-5),
    "C1b": dict(D=64, C=512, R=256, n=12, H1=512, ln_eps=1e-5),
    "C1c": dict(D=400, C=512, R=256, n=12, H1=512, ln_eps=1e-5),
    "C2a": dict(D=8, C=256, R=64, n=6, H1=256, ln_eps=1e-5),
    "C3": dict(D=2, C=256, R=64, n=6, H1=320, ln_eps=1e-6),
    "C3t": dict(D=2, C=256, R=64, n=6, H1=256, ln_eps=1e-5),
    "C4": dict(D=9, C=512, R=1024, n=12, H1=512, ln_eps=1e-5),
    "C5": dict(D=32, C=1024, R=256, n=24, H1=1024, ln_eps=1e-5),
}


def flops_fwd_per_sample(D, C, R, n, H1=None, **_):
    """F_fwd = n [2 D^2 + 4((dc+R) H1 + H1 C + C dt)]."""
    dc, dt, H1 = dims(D, C, R, H1)
    return n * (2 * D * D + 4 * ((dc + R) * H1 + H1 * C + C * dt))


KINK_MARGIN = 2e-6


def make_safe_inputs(sd, n, B, D, R, seed=1234, eps=1e-5, max_tries=400):
    """make_inputs with the first seed >= `seed` whose LeakyReLU inputs all stay KINK_MARGIN away
    from zero (see flow_oracle.min_abs_preactivation).  Returns (x, y, seed_used)."""
    from . import flow_oracle as fo
    sd64 = fo.cast_state_dict(sd, np.float64)
    for s in range(seed, seed + max_tries):
        x, y = make_inputs(B, D, R, seed=s)
        if fo.min_abs_preactivation(sd64, n, x.astype(np.float64), y.astype(np.float64), eps) > KINK_MARGIN:
            return x, y, s
    raise RuntimeError("no kink-free input seed found")


# ---------------------------------------------------------------------------- conditioning encoder (row f2)
# (input_size, rep_size) of the reference callers: GCBC obs+goal (torch_nfgcbc.py:120), kitchen obs 60 x window 10
# (torch_nfbc_kitchen.py:151), ant obs 29 x ... (SURVEY.md section 8 f2 quotes in = 4100).
ENCODER_CONFIGS = {
    "gcbc": dict(input_size=58, rep_size=64),
    "kitchen": dict(input_size=600, rep_size=1024),
    "ant": dict(input_size=4100, rep_size=256),
}


def make_encoder_state_dict(input_size: int, rep_size: int, seed: int = 0, hidden: int = 512, n_hidden: int = 4,
                            dtype=np.float32) -> "OrderedDict[str, np.ndarray]":
    """Weights with the reference GEncoder's key names (torch_fcnf.py:160-169): fc1..fcN, norm1..normN, fc_out."""
    rs = np.random.RandomState(seed)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def uni(shape, fan_in):
        k = 1.0 / math.sqrt(fan_in)
        return rs.uniform(-k, k, size=shape)

    fan = input_size
    for i in range(1, n_hidden + 1):
        sd[f"fc{i}.weight"] = uni((hidden, fan), fan)
        sd[f"fc{i}.bias"] = uni((hidden,), fan)
        sd[f"norm{i}.weight"] = 1.0 + 0.1 * rs.standard_normal(hidden)
        sd[f"norm{i}.bias"] = 0.1 * rs.standard_normal(hidden)
        fan = hidden
    sd["fc_out.weight"] = uni((rep_size, hidden), hidden)
    sd["fc_out.bias"] = uni((rep_size,), hidden)
    # the reference registers fc1..fcN, fc_out, then norm1..normN (torch_fcnf.py:160-169): same key order here
    order = [k for k in sd if k.startswith("fc")] + [k for k in sd if k.startswith("norm")]
    return OrderedDict((k, np.ascontiguousarray(sd[k], dtype=dtype)) for k in order)


def make_encoder_inputs(B: int, input_size: int, rep_size: int, seed: int = 77, dtype=np.float32):
    """observations g ~ N(0,1) and an upstream gradient for the representation."""
    rs = np.random.RandomState(seed)
    g = rs.standard_normal((B, input_size))
    dout = rs.standard_normal((B, rep_size))
    return np.ascontiguousarray(g, dtype=dtype), np.ascontiguousarray(dout, dtype=dtype)
