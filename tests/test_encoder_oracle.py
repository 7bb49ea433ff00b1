"""Pin the numpy encoder oracle against fixtures produced by the real reference
GEncoder (tests/golden/make_golden_encoder.py)."""
import os

import numpy as np
import pytest

from oracle import encoder_oracle as eo
from tests.helpers import grad_probe, probe_err, rel_err, tol_vs_arbiter
from tests.helpers_encoder import ENC_CASES, load_encoder_case


def test_encoder_fixtures_present():
    assert len(ENC_CASES) == 3


@pytest.mark.parametrize("path", ENC_CASES, ids=[os.path.basename(p)[:-4] for p in ENC_CASES])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_encoder_oracle_matches_reference(path, prec):
    cfg, sd, g, dout, gold = load_encoder_case(path)
    dt = np.float32 if prec == "f32" else np.float64
    sdc = {k: v.astype(dt) for k, v in sd.items()}
    res = eo.encoder_backward(sdc, g.astype(dt), dout.astype(dt))
    for name in ("out", "dg"):
        tol = tol_vs_arbiter(gold[f"f32/{name}"], gold[f"f64/{name}"]) if prec == "f32" else 1e-10
        assert rel_err(res[name], gold[f"f64/{name}"]) <= tol, (name, rel_err(res[name], gold[f"f64/{name}"]), tol)
    noise = [0.0]
    errs = {}
    for k, gr in res["grads"].items():
        errs[k] = probe_err(grad_probe(gr, k), gold[f"f64/gprobe/{k}"], gr.size)
        noise.append(probe_err(gold[f"f32/gprobe/{k}"], gold[f"f64/gprobe/{k}"], gr.size))
    tol = max(2e-5, 2.0 * max(noise)) if prec == "f32" else 1e-9
    for k, e in errs.items():
        assert e <= tol, (k, e, tol)


def test_encoder_gradient_matches_finite_differences():
    from oracle import synth
    sd = {k: v.astype(np.float64) for k, v in synth.make_encoder_state_dict(7, 5, seed=2, hidden=16, n_hidden=2).items()}
    g, dout = synth.make_encoder_inputs(3, 7, 5, seed=9, dtype=np.float64)
    res = eo.encoder_backward(sd, g, dout)
    f = lambda gg: float((eo.encoder_forward(sd, gg) * dout).sum())
    h = 1e-6
    for (r, c) in [(0, 0), (1, 3), (2, 6)]:
        gp, gm = g.copy(), g.copy()
        gp[r, c] += h
        gm[r, c] -= h
        assert abs((f(gp) - f(gm)) / (2 * h) - res["dg"][r, c]) < 1e-6
