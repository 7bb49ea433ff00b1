"""Pin the numpy oracle against fixtures produced by the real reference module
(tests/golden/make_golden.py imports /root/reference/il-and-cil/utils/torch_fcnf.py)."""
import os

import numpy as np
import pytest

from oracle import flow_oracle as fo
from tests.helpers import golden_files, grad_probe, load_case, probe_err, rel_err, tol_vs_arbiter

CASES = golden_files()


def test_fixtures_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_oracle_matches_reference(path, prec):
    cfg, sd, x, y, dz, dld, g = load_case(path)
    dt = np.float32 if prec == "f32" else np.float64
    sdc = fo.cast_state_dict(sd, dt)
    res = fo.flow_backward(sdc, cfg["n"], x.astype(dt), y.astype(dt), dz.astype(dt), dld.astype(dt))
    rev = fo.flow_reverse(sdc, cfg["n"], g[f"{prec}/z"].astype(dt), y.astype(dt))

    def check(name, ours):
        ref64 = g[f"f64/{name}"]
        tol = tol_vs_arbiter(g[f"f32/{name}"], ref64) if prec == "f32" else 1e-9
        if name == "rev" and prec == "f64":
            tol = 1e-7   # triangular solves: LAPACK vs torch ordering
        assert rel_err(ours, ref64) <= tol, (name, rel_err(ours, ref64), tol)

    check("z", res["z"])
    check("log_dets", res["log_dets"])
    check("outputs", np.stack(res["outputs"]))
    check("dx", res["dx"])
    check("dy", res["dy"])
    # reverse of the reference's own z (same precision) -- compared with the reference's reverse
    ref_rev64 = g["f64/rev"]
    tol = tol_vs_arbiter(g["f32/rev"], ref_rev64) if prec == "f32" else 1e-7
    if prec == "f32":
        # rev(z32) differs from rev(z64) by the forward error as well; compare like with like
        assert rel_err(rev, g["f32/rev"]) <= max(10 * tol, 1e-4)
    else:
        assert rel_err(rev, ref_rev64) <= tol
    # Gradient noise is judged globally: the reference's own fp32-vs-fp64 error varies a lot
    # from one parameter (or 32-entry probe) to the next, so use its maximum over all keys.
    errs, noise = {}, [0.0]
    for k, gr in res["grads"].items():
        if f"f64/grad/{k}" in g.files:
            errs[k] = rel_err(gr, g[f"f64/grad/{k}"])
            noise.append(rel_err(g[f"f32/grad/{k}"], g[f"f64/grad/{k}"]))
        else:
            errs[k] = probe_err(grad_probe(gr, k), g[f"f64/gprobe/{k}"], gr.size)
            noise.append(probe_err(g[f"f32/gprobe/{k}"], g[f"f64/gprobe/{k}"], gr.size))
    # fp32 gradient noise grows with depth (~1e-4 at n=24); the f64 run is the real pin
    tol = max(5e-5, 2.0 * max(noise)) if prec == "f32" else 1e-8
    for k, e in errs.items():
        assert e <= tol, (k, e, tol)


def test_logdet_is_log_abs_det_jacobian():
    """SURVEY.md section 4 invariant (ii): log_dets equals log|det d z/d x| on a tiny flow."""
    from oracle import synth
    D, C, R, n = 5, 16, 3, 2
    sd = fo.cast_state_dict(synth.make_state_dict(D, C, R, n, seed=3), np.float64)
    x, y = synth.make_inputs(1, D, R, seed=5, dtype=np.float64)
    z, _, ld = fo.flow_forward(sd, n, x, y)
    J = np.zeros((D, D))
    h = 1e-6
    for j in range(D):
        xp, xm = x.copy(), x.copy()
        xp[0, j] += h
        xm[0, j] -= h
        J[:, j] = (fo.flow_forward(sd, n, xp, y)[0][0] - fo.flow_forward(sd, n, xm, y)[0][0]) / (2 * h)
    assert abs(np.log(abs(np.linalg.det(J))) - ld[0]) < 1e-6


def test_reverse_inverts_forward_and_reverse_logdet():
    from oracle import synth
    D, C, R, n = 9, 32, 7, 3
    sd = fo.cast_state_dict(synth.make_state_dict(D, C, R, n, seed=4), np.float64)
    x, y = synth.make_inputs(11, D, R, seed=6, dtype=np.float64)
    z, _, ld = fo.flow_forward(sd, n, x, y)
    xr, ldr = fo.flow_reverse(sd, n, z, y, return_logdet=True)
    assert np.max(np.abs(xr - x)) < 1e-9
    assert np.max(np.abs(ldr + ld)) < 1e-9


@pytest.mark.parametrize("path", [p for p in CASES if "synth_C1a" in p or "synth_C2a" in p or "refinit_d9" in p],
                         ids=lambda p: os.path.basename(p)[:-4])
def test_torch_restatement_matches_reference(path):
    """oracle/torch_flow.py (the stock-PyTorch bar of tools/eager_bar.py) against the same fixtures, in float64."""
    import torch
    from oracle.torch_flow import TorchFlow
    cfg, sd, x, y, dz, dld, g = load_case(path)
    m = TorchFlow(cfg["D"], cfg["C"], cfg["R"], cfg["n"], first_hidden=cfg.get("H1"), ln_eps=cfg.get("ln_eps", 1e-5))
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    m = m.double()
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    yt = torch.tensor(y, dtype=torch.float64, requires_grad=True)
    z, outs, ld = m(xt, yt)
    ((z * torch.tensor(dz, dtype=torch.float64)).sum() + (ld * torch.tensor(dld, dtype=torch.float64)).sum()).backward()
    assert rel_err(z.detach().numpy(), g["f64/z"]) <= 1e-10
    assert rel_err(ld.detach().numpy(), g["f64/log_dets"]) <= 1e-10
    assert rel_err(torch.stack(outs).detach().numpy(), g["f64/outputs"]) <= 1e-10
    assert rel_err(xt.grad.numpy(), g["f64/dx"]) <= 1e-9
    assert rel_err(yt.grad.numpy(), g["f64/dy"]) <= 1e-9
    with torch.no_grad():
        rev = m.reverse(torch.tensor(g["f64/z"]), yt)
    assert rel_err(rev.numpy(), g["f64/rev"]) <= 1e-8
