"""GPU parity of the conditioning encoder through the C ABI: against the fixtures
of the real reference GEncoder and against the numpy oracle on ragged batches."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder_oracle as eo
from oracle import synth
from tests.helpers import grad_probe, probe_err, rel_err, tol_vs_arbiter
from tests.helpers_encoder import ENC_CASES, load_encoder_case

pytestmark = pytest.mark.gpu
PRECISIONS = ["fp32_simt", "fp32"]


def _build(cfg, sd, precision, **kw):
    import nf_b200
    m = nf_b200.GEncoder(cfg["input_size"], cfg["rep_size"], precision=precision, **kw)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    return m.to("cuda:0")


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("path", ENC_CASES, ids=[os.path.basename(p)[:-4] for p in ENC_CASES])
def test_encoder_golden_forward_backward(path, precision):
    cfg, sd, g, dout, gold = load_encoder_case(path)
    m = _build(cfg, sd, precision)
    gt = torch.tensor(g, device="cuda:0", requires_grad=True)
    out = m(gt)
    (out * torch.tensor(dout, device="cuda:0")).sum().backward()
    for name, ours in (("out", out.detach()), ("dg", gt.grad)):
        tol = tol_vs_arbiter(gold[f"f32/{name}"], gold[f"f64/{name}"])
        err = rel_err(ours.cpu().numpy(), gold[f"f64/{name}"])
        assert err <= tol, (name, err, tol)
    noise, errs = [0.0], {}
    for k, p in m.named_parameters():
        gr = p.grad.cpu().numpy()
        errs[k] = probe_err(grad_probe(gr, k), gold[f"f64/gprobe/{k}"], gr.size)
        noise.append(probe_err(gold[f"f32/gprobe/{k}"], gold[f"f64/gprobe/{k}"], gr.size))
    tol = max(2e-5, 2.0 * max(noise))
    for k, e in errs.items():
        assert e <= tol, (k, e, tol)
    # inference path (no stash) gives the same output (split-K partial sums are combined in arrival order, so the
    # two runs may differ in the last bit)
    with torch.no_grad():
        assert rel_err(m(gt.detach()).cpu().numpy(), out.detach().cpu().numpy()) <= 1e-6


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("B", [1, 50, 257])
@pytest.mark.parametrize("input_size,rep_size,hidden,n_hidden,eps", [(29, 16, 64, 2, 1e-5), (60, 40, 256, 3, 1e-6)])
def test_encoder_oracle_parity_ragged(B, input_size, rep_size, hidden, n_hidden, eps, precision):
    sd = synth.make_encoder_state_dict(input_size, rep_size, seed=5, hidden=hidden, n_hidden=n_hidden)
    g, dout = synth.make_encoder_inputs(B, input_size, rep_size, seed=11)
    m = _build(dict(input_size=input_size, rep_size=rep_size), sd, precision, hidden=hidden, n_hidden=n_hidden, ln_eps=eps)
    gt = torch.tensor(g, device="cuda:0", requires_grad=True)
    out = m(gt)
    (out * torch.tensor(dout, device="cuda:0")).sum().backward()
    sd64 = {k: v.astype(np.float64) for k, v in sd.items()}
    ref = eo.encoder_backward(sd64, g.astype(np.float64), dout.astype(np.float64), eps)
    assert rel_err(out.detach().cpu().numpy(), ref["out"]) <= 1e-5
    assert rel_err(gt.grad.cpu().numpy(), ref["dg"]) <= 1e-5
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), ref["grads"][k]) <= 2e-5, k


def test_encoder_feeds_the_flow_and_receives_dy():
    """The reference's training step (torch_nfbc_ant.py:267-272): y = encoder(obs); z, _, ld = flow(x, y); the flow's
    dy is the encoder's upstream gradient."""
    import nf_b200
    B, D, R = 64, 8, 64
    fsd = synth.make_state_dict(D, 128, R, 2, seed=1)
    esd = synth.make_encoder_state_dict(58, R, seed=2, hidden=128, n_hidden=2)
    flow = nf_b200.RealNVP(D, 128, R, 2, None, precision="fp32")
    flow.load_state_dict({k: torch.tensor(v) for k, v in fsd.items()})
    flow = flow.to("cuda:0")
    enc = _build(dict(input_size=58, rep_size=R), esd, "fp32", hidden=128, n_hidden=2)
    x, _ = synth.make_inputs(B, D, R)
    obs, _ = synth.make_encoder_inputs(B, 58, R)
    xt, ot = torch.tensor(x, device="cuda:0"), torch.tensor(obs, device="cuda:0")
    loss = flow.nll(xt, enc(ot))
    loss.backward()
    # oracle: chain the two numpy restatements in float64
    from oracle import flow_oracle as fo
    e64 = {k: v.astype(np.float64) for k, v in esd.items()}
    f64 = fo.cast_state_dict(fsd, np.float64)
    y = eo.encoder_forward(e64, obs.astype(np.float64))
    z, _, ld = fo.flow_forward(f64, 2, x.astype(np.float64), y)
    ref_loss = -np.mean(fo.std_normal_log_prob(z) + ld)
    assert abs(loss.item() - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    fb = fo.flow_backward(f64, 2, x.astype(np.float64), y, z / B, -np.ones(B) / B)
    eb = eo.encoder_backward(e64, obs.astype(np.float64), fb["dy"])
    for k, p in enc.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), eb["grads"][k]) <= 5e-5, k
