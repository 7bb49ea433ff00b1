"""Causal-transformer flow on the GPU against the float64 checker oracle/torch_tarflow.py
(PARITY UNPINNED upstream: the checker is validated by its own invariants in test_tarflow_cpu.py).  Tolerance: the fp32
contract of BASELINE.md section 3, err(ours, ref64) <= max(1e-5, 2 err(ref32, ref64)) with err = max|d| / max|ref64|."""
import copy

import numpy as np
import pytest
import torch

import nf_b200
from nf_b200 import _lib
from oracle import torch_tarflow as tt
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

CASES = {
    "gcbc": dict(A=8, C=256, R=64, hd=64, nb=4, L=4, clamp=0.0, B=256),       # torch_nfgcbc.py:74-82 defaults
    "small": dict(A=5, C=64, R=16, hd=32, nb=3, L=2, clamp=0.0, B=37),
    "clamp": dict(A=8, C=128, R=40, hd=64, nb=2, L=2, clamp=1.5, B=130),
    "long": dict(A=19, C=96, R=12, hd=32, nb=2, L=1, clamp=0.0, B=20),
    "one": dict(A=1, C=32, R=8, hd=32, nb=2, L=1, clamp=0.0, B=9),
}


def build(c, precision, seed=0):
    torch.manual_seed(seed)
    ref = tt.RealNVP(c["A"], c["C"], c["R"], c["hd"], c["nb"], c["L"], None, logscale_clamp=c["clamp"])
    tt.perturb_output_heads(ref, seed=seed + 1)
    ours = nf_b200.TransformerRealNVP(c["A"], c["C"], c["R"], c["hd"], c["nb"], c["L"], None, logscale_clamp=c["clamp"],
                                      precision=precision)
    ours.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(seed + 2)
    x = torch.rand(c["B"], c["A"], 1, generator=g) * 2 - 1 + 0.1 * torch.randn(c["B"], c["A"], 1, generator=g)
    y = torch.randn(c["B"], 1, c["R"], generator=g)
    dz = torch.randn(c["B"], c["A"], 1, generator=g)
    dld = torch.randn(c["B"], generator=g)
    return ref, ours.cuda(), x, y, dz, dld


def run_ref(ref, x, y, dz, dld, dtype):
    m = copy.deepcopy(ref).to(dtype)
    xx, yy = x.detach().clone().to(dtype).requires_grad_(True), y.detach().clone().to(dtype).requires_grad_(True)
    z, outs, ld = m(xx, yy)
    ((z * dz.to(dtype)).sum() + (ld * dld.to(dtype)).sum()).backward()
    grads = {k: p.grad.detach().numpy() for k, p in m.named_parameters()}
    with torch.no_grad():
        xr = m.reverse(z.detach(), yy.detach())
    return dict(z=z.detach().numpy(), outs=[o.detach().numpy() for o in outs], ld=ld.detach().numpy(),
                dx=xx.grad.numpy(), dy=yy.grad.numpy(), grads=grads, rev=xr.numpy())


@pytest.mark.parametrize("precision", ["fp32", "fp32_simt"])
@pytest.mark.parametrize("name", list(CASES))
def test_forward_backward_reverse_match_the_checker(name, precision):
    c = CASES[name]
    ref, ours, x, y, dz, dld = build(c, precision)
    r64 = run_ref(ref, x, y, dz, dld, torch.float64)
    r32 = run_ref(ref, x, y, dz, dld, torch.float32)
    xt, yt = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    for it in range(3):                      # eager, capture, replay
        ours.zero_grad(set_to_none=True)
        xt.grad = yt.grad = None
        z, outs, ld = ours(xt, yt)
        ((z * dz.cuda()).sum() + (ld * dld.cuda()).sum()).backward()
        with torch.no_grad():
            xr = ours.reverse(z.detach(), yt.detach())
            xr2, ldr = ours.reverse(z.detach(), yt.detach(), return_logdet=True)
        torch.cuda.synchronize()

        def chk(a, key, what, floor=1e-5):
            tol = max(floor, 2.0 * rel_err(r32[key], r64[key]))
            e = rel_err(a.detach().cpu().numpy(), r64[key])
            assert e <= tol, (name, precision, it, what, e, tol)
        chk(z, "z", "z"); chk(ld, "ld", "logdet"); chk(xt.grad, "dx", "dx"); chk(yt.grad, "dy", "dy")
        assert len(outs) == c["nb"] and torch.equal(outs[-1], z)
        for i, o in enumerate(outs):
            e = rel_err(o.detach().cpu().numpy(), r64["outs"][i])
            assert e <= max(1e-5, 2.0 * rel_err(r32["outs"][i], r64["outs"][i])), (name, "outputs", i, e)
        # the inverse, isolated: our reverse of our z against the float64 checker's reverse of the same z; the noise
        # floor is the float32 checker's reverse of that z (the autoregressive loop feeds every rounding error of
        # dimension t into dimensions t+1.., so the inverse is noisier than the forward in any float32 implementation)
        if it == 0:
            zc = z.detach().cpu()
            with torch.no_grad():
                want = copy.deepcopy(ref).double().reverse(zc.double(), y.detach().double()).numpy()
                noise = rel_err(ref.reverse(zc, y.detach()).numpy(), want)
            e = rel_err(xr.cpu().numpy(), want)
            assert e <= max(2e-5, 3.0 * noise), (name, precision, "reverse", e, noise)
            assert rel_err(xr.cpu().numpy(), x.detach().numpy()) <= max(5e-5, 6.0 * noise), (name, "round trip")
        assert torch.equal(xr, xr2)
        assert rel_err(ldr.cpu().numpy(), -r64["ld"]) <= max(2e-5, 4.0 * rel_err(r32["ld"], r64["ld"])), (name, "reverse logdet")
        for k, p in ours.named_parameters():
            g64, g32 = r64["grads"][k], r32["grads"][k]
            scale = max(float(np.abs(g64).max()), 1e-30)
            if scale < 1e-12:
                assert p.grad is None or float(p.grad.abs().max()) < 1e-6
                continue
            e = rel_err(p.grad.cpu().numpy(), g64)
            assert e <= max(5e-5, 3.0 * rel_err(g32, g64)), (name, precision, it, k, e)


def test_default_dispatch_runs_the_tensor_core_gemms():
    c = CASES["gcbc"]
    ref, ours, x, y, dz, dld = build(c, "fp32")
    _lib.lib.nf_graphs_enable(0)
    try:
        _lib.lib.nf_prof_enable(1)
        z, _, ld = ours(x.cuda().requires_grad_(True), y.cuda())
        (z.sum() + ld.sum()).backward()
        torch.cuda.synchronize()
        prof = _lib.prof_collect()
    finally:
        _lib.lib.nf_prof_enable(0)
        _lib.lib.nf_graphs_enable(1)
    assert prof["tc_gemm"][2] >= c["nb"] * c["L"] * 8 and prof["tc_wgrad"][2] >= c["nb"] * c["L"] * 4


def test_input_gradient_only_and_no_grad_paths():
    c = CASES["small"]
    ref, ours, x, y, dz, dld = build(c, "fp32")
    r64 = run_ref(ref, x, y, dz, dld, torch.float64)
    xt = x.cuda().requires_grad_(True)
    z, _, ld = ours(xt, y.cuda())
    (gx,) = torch.autograd.grad((z * dz.cuda()).sum() + (ld * dld.cuda()).sum(), [xt])   # the denoise step's pattern
    assert all(p.grad is None for p in ours.parameters())
    assert rel_err(gx.cpu().numpy(), r64["dx"]) <= 2e-5
    with torch.no_grad():
        z2, outs2, ld2 = ours(x.cuda(), y.cuda())
    assert torch.equal(z2, z.detach()) and torch.equal(ld2, ld.detach())


def test_rows_are_independent_and_batches_ragged():
    c = dict(CASES["clamp"], B=300)
    ref, ours, x, y, dz, dld = build(c, "fp32")
    with torch.no_grad():
        z, _, ld = ours(x.cuda(), y.cuda())
        z1, _, ld1 = ours(x[:129].cuda(), y[:129].cuda())
        xr = ours.reverse(z[:1], y[:1].cuda())
    assert rel_err(z1.cpu().numpy(), z[:129].cpu().numpy()) <= 2e-6 and rel_err(ld1.cpu().numpy(), ld[:129].cpu().numpy()) <= 2e-6
    assert rel_err(xr.cpu().numpy(), x[:1].numpy()) <= 5e-5


def test_persistent_inverse_kernel_agrees_with_the_per_token_path():
    """Evaluation-sized batches (B <= 128) run the autoregressive inverse as ONE cooperative kernel per block
    (csrc/tarflow_ar.cu); larger ones launch the per-token kernel sequence.  Same function: both against the float64
    checker, and against each other, at the reference script's shape with 20 environments (torch_evaluations.py:79-82)."""
    c = dict(CASES["gcbc"], B=20)
    ref, ours, x, y, dz, dld = build(c, "fp32")
    with torch.no_grad():
        z = ours(x.cuda(), y.cuda())[0]
        want = copy.deepcopy(ref).double().reverse(z.cpu().double(), y.double()).numpy()
        noise = rel_err(ref.reverse(z.cpu(), y).numpy(), want)
        res = {}
        for v in (-1, 0):
            _lib.set_option("tf_ar", v)
            try:
                for _ in range(3):      # eager, capture, replay
                    xr, ldr = ours.reverse(z, y.cuda(), return_logdet=True)
                res[v] = (xr.cpu().numpy(), ldr.cpu().numpy())
            finally:
                _lib.set_option("tf_ar", -1)
    for v in (-1, 0):
        assert rel_err(res[v][0], want) <= max(2e-5, 3.0 * noise), (v, rel_err(res[v][0], want), noise)
    assert rel_err(res[-1][0], res[0][0]) <= max(2e-5, 3.0 * noise)
    assert rel_err(res[-1][1], res[0][1]) <= 2e-5
    # the kernel count shows which path ran
    _lib.lib.nf_launch_count(1)
    with torch.no_grad():
        ours.reverse(z, y.cuda())
    assert _lib.lib.nf_launch_count(1) <= 4 * c["nb"] + 4
