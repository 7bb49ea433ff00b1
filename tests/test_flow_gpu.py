"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path called through the
nn.Module / C-ABI is compared with the golden fixtures generated from the real reference and
with the numpy oracle on seeded inputs.  Nothing here reads /root/reference."""
import ctypes
import os

import numpy as np
import pytest
import torch

import nf_b200
from nf_b200 import _lib
from oracle import flow_oracle as fo
from oracle import synth
from tests.helpers import golden_files, grad_probe, load_case, probe_err, rel_err, tol_vs_arbiter

pytestmark = pytest.mark.gpu
CASES = golden_files()
PRECISIONS = ["fp32_simt", "fp32"]


def dev():
    return torch.device("cuda:0")


def build_model(cfg, sd, precision, **kw):
    m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, precision=precision, **kw)
    m.load_state_dict({k: torch.tensor(np.asarray(v, dtype=np.float32)) for k, v in sd.items()})
    return m.to(dev())


def t(a):
    return torch.tensor(np.asarray(a, dtype=np.float32), device=dev())


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_golden_forward_reverse_backward(path, precision):
    cfg, sd, x, y, dz, dld, g = load_case(path)
    model = build_model(cfg, sd, precision)
    xt, yt = t(x).requires_grad_(True), t(y).requires_grad_(True)
    z, outputs, log_dets = model(xt, yt)
    assert len(outputs) == cfg["n"] and z.shape == xt.shape and log_dets.shape == (cfg["B"],)

    def check(name, ours, floor=1e-5):
        ref64 = g[f"f64/{name}"]
        tol = tol_vs_arbiter(g[f"f32/{name}"], ref64, floor)
        e = rel_err(ours.detach().cpu().numpy(), ref64)
        assert e <= tol, (name, e, tol)

    check("z", z)
    check("log_dets", log_dets)
    check("outputs", torch.stack(outputs))
    assert torch.equal(outputs[-1], z.detach())

    (z * t(dz)).sum().add((log_dets * t(dld)).sum()).backward()
    check("dx", xt.grad)
    check("dy", yt.grad)
    errs, noise = {}, [0.0]
    for k, p in model.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        gr = p.grad.detach().cpu().numpy()
        if f"f64/grad/{k}" in g.files:
            errs[k] = rel_err(gr, g[f"f64/grad/{k}"])
            noise.append(rel_err(g[f"f32/grad/{k}"], g[f"f64/grad/{k}"]))
        else:
            errs[k] = probe_err(grad_probe(gr, k), g[f"f64/gprobe/{k}"], gr.size)
            noise.append(probe_err(g[f"f32/gprobe/{k}"], g[f"f64/gprobe/{k}"], gr.size))
    tol = max(5e-5, 2.0 * max(noise))
    bad = {k: e for k, e in errs.items() if e > tol}
    assert not bad, (sorted(bad.items(), key=lambda kv: -kv[1])[:5], tol)

    # reverse of the reference's own z, against the reference's own reverse
    with torch.no_grad():
        rev = model.reverse(t(g["f32/z"]), t(y))
    tol = max(10 * tol_vs_arbiter(g["f32/rev"], g["f64/rev"]), 1e-4)
    assert rel_err(rev.cpu().numpy(), g["f32/rev"]) <= tol
    # NLL of torch_nfbc_ant.py:269 through the log_prob extra
    with torch.no_grad():
        nll = -model.log_prob(t(x), t(y)).mean().item()
    assert abs(nll - float(g["f64/nll"])) <= 1e-5 * max(1.0, abs(float(g["f64/nll"]))) + 2 * abs(
        float(g["f32/nll"]) - float(g["f64/nll"]))


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("B", [1, 20, 50, 257])
@pytest.mark.parametrize("D,C,R,n,H1,eps", [(9, 64, 24, 3, None, 1e-5), (2, 32, 8, 2, 40, 1e-6), (8, 48, 16, 2, None, 1e-5)])
def test_oracle_parity_ragged_batches(B, D, C, R, n, H1, eps, precision):
    """Odd D (9 -> 5/4 split), batch sizes that are not tile multiples, the flax twin's widths
    (first hidden C+R, eps 1e-6) -- against the numpy oracle in float64."""
    sd = synth.make_state_dict(D, C, R, n, seed=11, H1=H1)
    x, y, _ = synth.make_safe_inputs(sd, n, B, D, R, seed=100 * B, eps=eps)   # away from LeakyReLU kinks
    dz, dld = synth.make_cotangents(B, D, seed=B + 1)
    cfg = dict(D=D, C=C, R=R, n=n, B=B)
    model = build_model(cfg, sd, precision, first_hidden=H1, ln_eps=eps)
    sd64 = fo.cast_state_dict(sd, np.float64)
    ref = fo.flow_backward(sd64, n, x.astype(np.float64), y.astype(np.float64), dz.astype(np.float64),
                           dld.astype(np.float64), eps=eps)
    xt, yt = t(x).requires_grad_(True), t(y).requires_grad_(True)
    z, outputs, ld = model(xt, yt)
    (z * t(dz)).sum().add((ld * t(dld)).sum()).backward()
    tol = 2e-5
    assert rel_err(z.detach().cpu().numpy(), ref["z"]) <= tol
    assert rel_err(ld.detach().cpu().numpy(), ref["log_dets"]) <= tol
    assert rel_err(xt.grad.cpu().numpy(), ref["dx"]) <= tol
    assert rel_err(yt.grad.cpu().numpy(), ref["dy"]) <= tol
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), ref["grads"][k]) <= 5e-5, k
    with torch.no_grad():
        seq, rld = model.reverse(z.detach(), yt.detach(), return_sequence=True, return_logdet=True)
    ref_seq, ref_rld = fo.flow_reverse(sd64, n, ref["z"], y.astype(np.float64), eps=eps, return_sequence=True,
                                       return_logdet=True)
    assert len(seq) == n + 1
    for a, b in zip(seq, ref_seq):
        assert rel_err(a.cpu().numpy(), b) <= 1e-4
    assert rel_err(rld.cpu().numpy(), ref_rld) <= 1e-4
    # inverse(forward(x)) == x and reverse log-det == -forward log-det
    assert rel_err(seq[-1].cpu().numpy(), x) <= 1e-4
    assert rel_err(rld.cpu().numpy(), -ref["log_dets"]) <= 1e-4


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("cfgname,B", [("C2a", 4096), ("C4", 8192), ("C3", 16384)])
def test_full_size_properties(cfgname, B, precision):
    """BASELINE.json sizes: size-independent properties (round trip, log-det antisymmetry,
    shard-invariance of outputs and of data-parallel gradients)."""
    cfg = dict(synth.CONFIGS[cfgname])
    eps, H1 = cfg.pop("ln_eps"), cfg.pop("H1")
    sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
    model = build_model(dict(cfg, B=B), sd, precision, first_hidden=H1, ln_eps=eps)
    x, y = synth.make_inputs(B, cfg["D"], cfg["R"], seed=99)
    xt, yt = t(x), t(y)
    with torch.no_grad():
        z, outs, ld = model(xt, yt)
        xr, rld = model.reverse(z, yt, return_logdet=True)
    assert torch.isfinite(z).all() and torch.isfinite(ld).all()
    # the reference's own fp32 round trip is 3e-5 ... 1.7e-2
    assert rel_err(xr.cpu().numpy(), x) <= 2e-3
    assert rel_err(rld.cpu().numpy(), -ld.cpu().numpy()) <= 2e-5
    # rows are independent: a shard gives the same rows as the full batch
    h = B // 2 + 3
    with torch.no_grad():
        z2, _, ld2 = model(xt[h:], yt[h:])   # row slice: contiguous but not 16-byte aligned
    assert rel_err(z2.cpu().numpy(), z[h:].cpu().numpy()) <= 1e-6
    assert rel_err(ld2.cpu().numpy(), ld[h:].cpu().numpy()) <= 1e-6
    if cfgname == "C4":
        return   # sampling-only config
    # data parallel invariant: mean of shard gradients == full-batch gradient of the mean loss
    def grads(xs, ys, scale):
        for p in model.parameters():
            p.grad = None
        (-model.log_prob(xs, ys).sum() * scale).backward()
        return torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad])
    full = grads(xt, yt, 1.0 / B)
    half = B // 2
    parts = 0.5 * (grads(xt[:half].contiguous(), yt[:half].contiguous(), 1.0 / half)
                   + grads(xt[half:].contiguous(), yt[half:].contiguous(), 1.0 / half))
    assert rel_err(parts.cpu().numpy(), full.cpu().numpy()) <= 2e-5


def test_denoise_step_input_gradient_only():
    """torch_nfbc_ant.py:212-217: autograd.grad of the log-prob w.r.t. the action only."""
    cfg = dict(D=8, C=64, R=32, n=3, B=20)
    sd = synth.make_state_dict(8, 64, 32, 3, seed=2)
    model = build_model(cfg, sd, PRECISIONS[0])
    for p in model.parameters():
        p.requires_grad_(False)
    x, y, _ = synth.make_safe_inputs(sd, 3, 20, 8, 32, seed=3)
    action = t(x).requires_grad_(True)
    z, _, ld = model(action, t(y))
    logprob = -0.5 * (z * z).sum(1) + ld
    (grad,) = torch.autograd.grad(logprob.sum(), [action])
    sd64 = fo.cast_state_dict(sd, np.float64)
    z64, _, _ = fo.flow_forward(sd64, 3, x.astype(np.float64), y.astype(np.float64))
    ref = fo.flow_backward(sd64, 3, x.astype(np.float64), y.astype(np.float64), -z64, np.ones(20))
    assert rel_err(grad.cpu().numpy(), ref["dx"]) <= 2e-5


@pytest.mark.parametrize("precision", PRECISIONS)
def test_denoise_call_matches_the_reference_sequence(precision):
    """RealNVP.denoise == the reference's clone / forward / autograd.grad / in-place update (torch_nfbc_ant.py:212-217),
    with trainable parameters left alone (no gradients appear on them)."""
    cfg = dict(D=9, C=64, R=24, n=3, B=50)
    sd = synth.make_state_dict(9, 64, 24, 3, seed=4)
    model = build_model(cfg, sd, precision)
    x, y, _ = synth.make_safe_inputs(sd, 3, 50, 9, 24, seed=8)
    noise_std = 0.1
    out = model.denoise(t(x), t(y), noise_std)
    assert all(p.grad is None for p in model.parameters())
    sd64 = fo.cast_state_dict(sd, np.float64)
    z64, _, _ = fo.flow_forward(sd64, 3, x.astype(np.float64), y.astype(np.float64))
    ref = fo.flow_backward(sd64, 3, x.astype(np.float64), y.astype(np.float64), -z64, np.ones(50))
    want = x.astype(np.float64) + noise_std ** 2 * ref["dx"]
    assert rel_err(out.cpu().numpy(), want) <= 1e-5
    # the same through autograd, as the reference script writes it
    action = t(x).clone().requires_grad_(True)
    z, _, ld = model(action, t(y))
    (grad,) = torch.autograd.grad((-0.5 * (z * z).sum(1) + ld).sum(), [action])
    assert rel_err(out.cpu().numpy(), (action.detach() + noise_std ** 2 * grad).cpu().numpy()) <= 1e-6
    # sample + denoise stays close to the plain sample (a small step)
    torch.manual_seed(0)
    a = model.sample_denoised(t(y), noise_std)
    assert a.shape == (50, 9) and torch.isfinite(a).all()


def test_empty_batch():
    """B = 0 (SURVEY section 8c edge cases): every call returns empty tensors and zero parameter gradients."""
    cfg = dict(D=8, C=32, R=16, n=2, B=0)
    sd = synth.make_state_dict(8, 32, 16, 2, seed=1)
    model = build_model(cfg, sd, PRECISIONS[1])
    x = torch.empty((0, 8), device="cuda:0", requires_grad=True)
    y = torch.empty((0, 16), device="cuda:0")
    z, outs, ld = model(x, y)
    assert z.shape == (0, 8) and ld.shape == (0,) and len(outs) == 2 and outs[0].shape == (0, 8)
    (z.sum() + ld.sum()).backward()
    assert x.grad.shape == (0, 8)
    for p in model.parameters():
        if p.requires_grad:
            assert p.grad is not None and float(p.grad.abs().max()) == 0.0
    assert model.reverse(x.detach(), y).shape == (0, 8)
    assert model.denoise(x.detach(), y, 0.1).shape == (0, 8)


def test_standalone_affine_and_plu_kernels():
    lib = _lib.lib
    rs = np.random.RandomState(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for D, B in [(8, 1000), (9, 257), (400, 64), (2, 5000)]:
        dc, dt = (D + 1) // 2, D // 2
        xp, s, tt = rs.randn(B, D), 0.3 * rs.randn(B, dt), rs.randn(B, dt)
        bs, bt = 0.1 * rs.randn(dt), 0.1 * rs.randn(dt)
        ld0 = rs.randn(B)
        z = torch.empty(B, D, device=dev())
        ld = t(ld0)
        sout = torch.empty(B, dt, device=dev())
        xp_d, s_d, t_d, bs_d, bt_d = t(xp), t(s), t(tt), t(bs), t(bt)   # keep the device buffers alive
        _lib.check(lib.nf_affine_logdet_fwd(xp_d.data_ptr(), s_d.data_ptr(), t_d.data_ptr(), bs_d.data_ptr(),
                                            bt_d.data_ptr(), 0.25, B, D, z.data_ptr(), ld.data_ptr(),
                                            sout.data_ptr(), st), "affine_fwd")
        sf, tf = s + bs, tt + bt
        zref = np.concatenate([xp[:, :dc], (xp[:, dc:] - tf) * np.exp(-sf)], axis=1)
        assert rel_err(z.cpu().numpy(), zref) <= 1e-6
        assert rel_err(ld.cpu().numpy(), ld0 + 0.25 - sf.sum(1)) <= 1e-6
        assert rel_err(sout.cpu().numpy(), sf) <= 1e-6
        # inverse undoes forward
        back = torch.empty(B, D, device=dev())
        ld2 = ld.clone()
        _lib.check(lib.nf_affine_logdet_inv(z.data_ptr(), s_d.data_ptr(), t_d.data_ptr(), bs_d.data_ptr(),
                                            bt_d.data_ptr(), 0.25, B, D, back.data_ptr(), ld2.data_ptr(), st),
                   "affine_inv")
        assert rel_err(back.cpu().numpy(), xp) <= 1e-5
        assert rel_err(ld2.cpu().numpy(), ld0) <= 1e-5
        # backward
        dzv, dldv = rs.randn(B, D), rs.randn(B)
        dxp = torch.empty(B, D, device=dev()); ds = torch.empty(B, dt, device=dev()); dtt = torch.empty(B, dt, device=dev())
        dz_d, dld_d = t(dzv), t(dldv)
        _lib.check(lib.nf_affine_logdet_bwd(dz_d.data_ptr(), z.data_ptr(), sout.data_ptr(), dld_d.data_ptr(), B, D,
                                            dxp.data_ptr(), ds.data_ptr(), dtt.data_ptr(), st), "affine_bwd")
        e = np.exp(-sf)
        assert rel_err(dxp.cpu().numpy(), np.concatenate([dzv[:, :dc], dzv[:, dc:] * e], 1)) <= 1e-6
        assert rel_err(ds.cpu().numpy(), -dzv[:, dc:] * zref[:, dc:] - dldv[:, None]) <= 1e-5
        assert rel_err(dtt.cpu().numpy(), -dzv[:, dc:] * e) <= 1e-6
    for D in (2, 9, 64, 400):
        sd = synth.make_state_dict(D, 4, 2, 1, seed=D)
        W = torch.empty(D, D, device=dev()); Wi = torch.empty(D, D, device=dev()); ldt = torch.empty(1, device=dev())
        ws = torch.empty(64 * D * D + 4096, dtype=torch.float32, device=dev())
        pd = [t(sd["blocks.0.l." + k]) for k in ("s", "U", "L", "P", "P_inv")]
        _lib.check(lib.nf_plu_build(pd[0].data_ptr(), pd[1].data_ptr(), pd[2].data_ptr(), pd[3].data_ptr(),
                                    pd[4].data_ptr(), D, W.data_ptr(), Wi.data_ptr(),
                                    ldt.data_ptr(), ws.data_ptr(), ws.numel() * 4, st), "plu_build")
        sd64 = fo.cast_state_dict(sd, np.float64)
        assert rel_err(W.cpu().numpy(), fo.plu_weight(sd64, 0)[0]) <= 1e-5
        assert rel_err(Wi.cpu().numpy(), fo.plu_weight_inv(sd64, 0)) <= 1e-4
        assert abs(ldt.item() - fo.plu_logdet(sd64, 0)) <= 1e-4


def test_abi_argument_errors_on_device():
    lib = _lib.lib
    d = _lib.make_desc(8, 32, 32, 4, 2, 1e-5, 0)
    buf = torch.zeros(1 << 16, device=dev())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # misaligned pointer
    rc = lib.nf_flow_forward(ctypes.byref(d), buf.data_ptr() + 4, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4,
                             buf.data_ptr(), buf.data_ptr(), None, None, buf.data_ptr(), buf.numel() * 4, st)
    assert rc == -2 and b"aligned" in lib.nf_last_error()
    # workspace too small
    rc = lib.nf_flow_forward(ctypes.byref(d), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4096,
                             buf.data_ptr(), buf.data_ptr(), None, None, buf.data_ptr(), 64, st)
    assert rc == -3
    # empty batch is a no-op
    rc = lib.nf_flow_forward(ctypes.byref(d), buf.data_ptr(), buf.data_ptr(), None, None, 0, None, None, None, None,
                             None, 0, st)
    assert rc == 0
    m = nf_b200.RealNVP(8, 32, 4, 2, None).to(dev())
    z, outs, ld = m(torch.zeros(0, 8, device=dev()), torch.zeros(0, 4, device=dev()))
    assert z.shape == (0, 8) and ld.shape == (0,)


def test_fresh_model_is_identity_coupling_and_state_dict_roundtrip(tmp_path):
    """Zero-initialised heads (torch_fcnf.py:79-82): a fresh model has s = t = 0, so
    log_dets == sum_i sum log|s_plu| for every row; checkpoints keep the reference's keys
    (il-and-cil/utils/torch_utils.py:5-16)."""
    torch.manual_seed(0)
    m = nf_b200.RealNVP(8, 64, 16, 4, None).to(dev())
    x, y = synth.make_inputs(33, 8, 16)
    z, outs, ld = m(t(x), t(y))
    expect = sum(float(torch.log(torch.abs(b.l.s)).sum()) for b in m.blocks)
    assert torch.allclose(ld, torch.full_like(ld, expect), atol=1e-5)
    path = os.path.join(tmp_path, "params.pth")
    torch.save({"model": m.state_dict()}, path)
    m2 = nf_b200.RealNVP(8, 64, 16, 4, None).to(dev())
    m2.load_state_dict(torch.load(path)["model"])
    z2, _, _ = m2(t(x), t(y))
    assert torch.equal(z, z2)
    # one optimiser step changes the output (prepared tables are refreshed)
    opt = torch.optim.AdamW(m2.parameters(), lr=1e-2)
    loss = -m2.log_prob(t(x), t(y)).mean()
    loss.backward()
    opt.step()
    z3, _, _ = m2(t(x), t(y))
    assert not torch.equal(z2, z3)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_fast_param_grads_match_autograd_path_and_accumulate(precision):
    """fast_param_grads (flat gradient arena + p.grad views) gives the gradients the plain autograd
    path returns, accumulates like autograd when gradients are already present, and respects
    frozen parameters."""
    D, C, R, n, B = 8, 64, 16, 3, 130
    sd = synth.make_state_dict(D, C, R, n, seed=21)
    x, y, _ = synth.make_safe_inputs(sd, n, B, D, R, seed=22)
    cfg = dict(D=D, C=C, R=R, n=n, B=B)
    slow = build_model(cfg, sd, precision, fast_param_grads=False)
    fast = build_model(cfg, sd, precision, fast_param_grads=True)

    def run(m):
        xt = t(x).requires_grad_(True)
        (-m.log_prob(xt, t(y)).mean()).backward()
        return xt.grad

    gx_s, gx_f = run(slow), run(fast)
    tol = 0.0 if precision == "fp32_simt" else 1e-6      # atomics order differs between calls on the tensor path
    assert rel_err(gx_f.cpu().numpy(), gx_s.cpu().numpy()) <= 1e-6
    ref = {k: p.grad.clone() for k, p in slow.named_parameters() if p.requires_grad}
    for k, p in fast.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and p.grad.shape == p.shape
            assert rel_err(p.grad.cpu().numpy(), ref[k].cpu().numpy()) <= max(tol, 2e-6), k
        else:
            assert p.grad is None
    # second backward without zero_grad accumulates (autograd semantics), in both modes
    run(slow); run(fast)
    for k, p in fast.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), 2 * ref[k].cpu().numpy()) <= 5e-6, k
            assert rel_err(dict(slow.named_parameters())[k].grad.cpu().numpy(), 2 * ref[k].cpu().numpy()) <= 5e-6, k
    # zero_grad(set_to_none=True) then a third backward: back to a single gradient
    fast.zero_grad(set_to_none=True)
    run(fast)
    for k, p in fast.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), ref[k].cpu().numpy()) <= 5e-6, k
    # a foreign gradient tensor (not our view) is accumulated into, not replaced
    fast.zero_grad(set_to_none=True)
    w = fast.blocks[0].t[0].weight
    w.grad = torch.ones_like(w)
    run(fast)
    assert rel_err((w.grad - 1).cpu().numpy(), ref["blocks.0.t.0.weight"].cpu().numpy()) <= 5e-6
    # frozen parameters get no gradient and do not block the others
    fast.zero_grad(set_to_none=True)
    fast.blocks[0].t[0].weight.requires_grad_(False)
    run(fast)
    assert fast.blocks[0].t[0].weight.grad is None
    assert rel_err(fast.blocks[1].s[3].weight.grad.cpu().numpy(), ref["blocks.1.s.3.weight"].cpu().numpy()) <= 5e-6
    # optimizer step through the views changes the arena the kernels read
    opt = torch.optim.SGD([p for p in fast.parameters() if p.requires_grad], lr=0.1)
    before = fast._flat.clone()
    opt.step()
    assert not torch.equal(before, fast._flat)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_launch_graph_replay_is_bit_identical(precision):
    """The library replays a cached CUDA graph when the same (descriptor, B, pointers) come back: the
    replayed forward / inverse give bit-identical results to the eager first call, also after the
    parameters changed in place (the graph reads the arena, it does not bake values in)."""
    D, C, R, n, B = 9, 64, 24, 3, 200
    sd = synth.make_state_dict(D, C, R, n, seed=31)
    x, y, _ = synth.make_safe_inputs(sd, n, B, D, R, seed=32)
    model = build_model(dict(D=D, C=C, R=R, n=n, B=B), sd, precision)
    xt, yt = t(x), t(y)
    r0 = _lib.graph_stats()
    outs = []
    with torch.no_grad():
        for _ in range(4):
            z, _, ld = model(xt, yt)
            outs.append((z.clone(), ld.clone()))
            del z, ld                                   # same addresses come back from the caching allocator
    r1 = _lib.graph_stats()
    assert r1[0] > r0[0], "no graph replay happened"
    for z, ld in outs[1:]:
        assert torch.equal(z, outs[0][0]) and torch.equal(ld, outs[0][1])
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad:
                p.mul_(1.01)
        z2, _, ld2 = model(xt, yt)
        _lib.lib.nf_graphs_enable(0)
        try:
            z3, _, ld3 = model(xt, yt)
        finally:
            _lib.lib.nf_graphs_enable(1)
    assert not torch.equal(z2, outs[0][0])
    assert torch.equal(z2, z3) and torch.equal(ld2, ld3)


OPTION_SETTINGS = [("pdl", 0), ("streams", 0), ("tma_store", 0), ("fused_ln", 0), ("fused_ln", 1), ("fused_tail", 1),
                   ("nacc", 1), ("bn256", 2), ("bn256", 1), ("ts", 0), ("presplit", 0),
                   ("stack", 0), ("stack", 1), ("stack_nsg", 1), ("ride", 0), ("defer_wgrad", 0), ("defer_wgrad", 1), ("mcast", 0), ("mcast", 1),
                   ("prep_fused", 0), ("defer_dy", 0), ("defer_cols", 0), ("defer_skinny", 1)]


@pytest.mark.parametrize("name,value", OPTION_SETTINGS, ids=[f"{n}={v}" for n, v in OPTION_SETTINGS])
def test_options_are_equivalent(name, value):
    """Every execution option (PDL, side streams, TMA-store epilogue, fused LayerNorm / tail epilogues,
    accumulator rotation, tile width) computes the same function: forward, inverse and all gradients agree
    with the default configuration to fp32 rounding, eagerly and through the replayed launch graph."""
    D, C, R, n, B = 8, 256, 64, 2, 300
    sd = synth.make_state_dict(D, C, R, n, seed=41)
    x, y, _ = synth.make_safe_inputs(sd, n, B, D, R, seed=42)
    dz, dld = synth.make_cotangents(B, D, seed=43)

    def run():
        model = build_model(dict(D=D, C=C, R=R, n=n, B=B), sd, "fp32")
        outs = None
        for _ in range(3):          # eager, capture, replay
            model.zero_grad(set_to_none=True)
            xt, yt = t(x).requires_grad_(True), t(y).requires_grad_(True)
            z, _, ld = model(xt, yt)
            (z * t(dz)).sum().add((ld * t(dld)).sum()).backward()
            with torch.no_grad():
                xr = model.reverse(z.detach(), t(y))
            g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad])
            cur = [v.detach().cpu().numpy().copy() for v in (z, ld, xt.grad, yt.grad, g, xr)]
            if outs is not None:
                for a, b, nm in zip(outs, cur, ("z", "ld", "dx", "dy", "dparams", "rev")):
                    assert rel_err(b, a) <= 2e-6, ("replay differs from the eager run", nm)
            outs = cur
        return outs

    default = _lib.get_option(name)
    ref = run()                              # library defaults (forward / inverse: the persistent whole-stack kernel)
    stack_default = _lib.get_option("stack")
    if not name.startswith("stack"):
        _lib.set_option("stack", 0)          # the per-layer kernels are what these options change
    _lib.set_option(name, value)
    try:
        assert _lib.get_option(name) == value
        got = run()
    finally:
        _lib.set_option(name, default)
        _lib.set_option("stack", stack_default)
    tol = 1e-5 if name == "nacc" else 3e-6   # one accumulator instead of three: the truncation noise itself changes
    for a, b, nm in zip(ref, got, ("z", "ld", "dx", "dy", "dparams", "rev")):
        assert rel_err(b, a) <= tol, (name, value, nm, rel_err(b, a))


@pytest.mark.parametrize("shape", [dict(D=8, C=256, R=64, n=3, H1=256), dict(D=2, C=256, R=64, n=2, H1=320),
                                   dict(D=9, C=128, R=37 * 4, n=2, H1=128), dict(D=32, C=64, R=16, n=2, H1=96)],
                         ids=["C2a-like", "C3-like", "odd-D", "D32"])
def test_one_kernel_prepare_writes_the_same_tables(shape):
    """nf_flow_prepare(NF_PREP_FWD) for D <= 32 is one kernel that writes every derived table with its hi / lo copies
    directly (k_prepare_fwd); the graph of separate kernels it replaces (option prep_fused = 0) must produce the same
    bytes -- same products, same summation order."""
    D, C, R, n, H1 = shape["D"], shape["C"], shape["R"], shape["n"], shape["H1"]
    sd = synth.make_state_dict(D, C, R, n, seed=77, H1=H1)
    model = build_model(dict(D=D, C=C, R=R, n=n, B=64, H1=H1), sd, "fp32", first_hidden=H1)
    x, y = synth.make_inputs(8, D, R, seed=3)
    with torch.no_grad():
        model(t(x), t(y))                    # allocates the table buffer
    tables = []
    default = _lib.get_option("prep_fused")
    try:
        for v in (1, 0):
            _lib.set_option("prep_fused", v)
            model._packed.zero_()
            model._packed_key = None
            model._prepared(_lib.NF_PREP_FWD, model._flat.device)
            torch.cuda.synchronize()
            tables.append(model._packed.clone())
    finally:
        _lib.set_option("prep_fused", default)
    a, b = (v.view(torch.float32) for v in tables)
    assert float(a.abs().max()) > 0
    assert torch.equal(a, b), int((a != b).sum())


def test_fused_nll_and_flat_adamw_match_the_torch_training_step():
    """model.nll() (fused loss + cotangents) and FlatAdamW (one kernel over the arena)
    reproduce `loss = -(log N(z) + logdets).mean(); loss.backward(); torch.optim.AdamW.step()` of
    torch_nfbc_ant.py:269-273 for a few steps, with an extra (encoder-like) parameter in the same optimizer and
    the reference-style cosine schedule driving the learning rate."""
    D, C, R, n, B = 8, 64, 16, 3, 96
    sd = synth.make_state_dict(D, C, R, n, seed=51)
    cfg = dict(D=D, C=C, R=R, n=n, B=B)
    a = build_model(cfg, sd, "fp32_simt", fast_param_grads=False)      # plain torch training step
    b = build_model(cfg, sd, "fp32_simt")                              # fused loss + fused optimizer
    torch.manual_seed(0)
    enc_a = torch.nn.Linear(5, R).to(dev())
    enc_b = torch.nn.Linear(5, R).to(dev())
    enc_b.load_state_dict(enc_a.state_dict())
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-2)
    opt_a = torch.optim.AdamW(list(a.parameters()) + list(enc_a.parameters()), **kw)
    opt_b = nf_b200.FlatAdamW(list(b.parameters()) + list(enc_b.parameters()), flows=[b], **kw)
    sch_a = nf_b200.CosineLRSchedule(opt_a, warmup_steps=2, total_steps=6, min_lr=1e-4, max_lr=3e-3)
    sch_b = nf_b200.CosineLRSchedule(opt_b, warmup_steps=2, total_steps=6, min_lr=1e-4, max_lr=3e-3)
    LOG2PI = float(np.log(2 * np.pi))
    for it in range(4):
        x, _ = synth.make_inputs(B, D, R, seed=60 + it)
        obs = torch.randn(B, 5, generator=torch.Generator().manual_seed(it)).to(dev())
        opt_a.zero_grad(set_to_none=True)
        z, _, ld = a(t(x), enc_a(obs))
        loss_a = -((-0.5 * (z * z).sum(1) - 0.5 * D * LOG2PI) + ld).mean()
        loss_a.backward()
        opt_a.step(); sch_a.step()
        opt_b.zero_grad(set_to_none=True)
        loss_b = b.nll(t(x), enc_b(obs))
        loss_b.backward()
        opt_b.step(); sch_b.step()
        assert abs(loss_a.item() - loss_b.item()) <= 2e-6 * max(1.0, abs(loss_a.item())), it
        assert opt_a.param_groups[0]["lr"] == opt_b.param_groups[0]["lr"]
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    for k in pa:
        assert rel_err(pb[k].detach().cpu().numpy(), pa[k].detach().cpu().numpy()) <= 5e-6, k
    for k, v in enc_a.state_dict().items():
        assert rel_err(enc_b.state_dict()[k].cpu().numpy(), v.cpu().numpy()) <= 5e-6, k
    # frozen PLU permutations are untouched by the optimizer (they carry weight decay otherwise)
    assert torch.equal(b.blocks[0].l.P, t(sd["blocks.0.l.P"]))


def test_overlapped_data_parallel_backward_matches_single_call(tmp_path):
    """The data-parallel backward runs block range by block range (nf_flow_backward_range) with an asynchronous
    NCCL all-reduce per gradient bucket.  On a one-rank NCCL group the average is the identity, so the result must
    equal the single-call backward: gradients w.r.t. x, y and every parameter, for several bucket sizes."""
    import torch.distributed as dist
    D, C, R, n, B = 8, 64, 16, 5, 150
    sd = synth.make_state_dict(D, C, R, n, seed=71)
    x, y, _ = synth.make_safe_inputs(sd, n, B, D, R, seed=72)
    cfg = dict(D=D, C=C, R=R, n=n, B=B)

    def grads(model):
        model.zero_grad(set_to_none=True)
        xt, yt = t(x).requires_grad_(True), t(y).requires_grad_(True)
        (-model.log_prob(xt, yt).mean()).backward()
        g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad])
        return [v.detach().cpu().numpy().copy() for v in (xt.grad, yt.grad, g)]

    ref = grads(build_model(cfg, sd, "fp32"))
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1,
                                device_id=dev())
    try:
        for bucket in (1, 2, 5):
            model = build_model(cfg, sd, "fp32")
            model.enable_data_parallel()
            model._dp_world = 2            # force the bucketed path; the one-rank average leaves values unchanged
            model.dp_bucket_blocks = bucket
            for _ in range(2):             # eager, then captured graphs per range
                got = grads(model)
                for a, b, nm in zip(ref, got, ("dx", "dy", "dparams")):
                    assert rel_err(b, a) <= 3e-6, (bucket, nm, rel_err(b, a))
    finally:
        if created:
            dist.destroy_process_group()


def test_denoise_step_leaves_unfrozen_parameters_alone():
    """The reference's evaluation step (torch_nfbc_ant.py:212-217) calls autograd.grad w.r.t. the action while the
    parameters still require grad: no p.grad may appear (a following optimizer step would otherwise train on
    evaluation gradients) and no weight-gradient kernel may run."""
    cfg = dict(D=8, C=64, R=32, n=3, B=20)
    sd = synth.make_state_dict(8, 64, 32, 3, seed=2)
    for fast in (True, False):
        model = build_model(cfg, sd, "fp32", fast_param_grads=fast)
        assert all(p.requires_grad for k, p in model.named_parameters() if not k.endswith(("P", "P_inv")))
        x, y, _ = synth.make_safe_inputs(sd, 3, 20, 8, 32, seed=3)
        action = t(x).requires_grad_(True)
        z, _, ld = model(action, t(y))
        logprob = -0.5 * (z * z).sum(1) + ld
        _lib.lib.nf_prof_enable(1)
        _lib.prof_collect()
        (grad,) = torch.autograd.grad(logprob.sum(), [action])
        torch.cuda.synchronize()
        prof = _lib.prof_collect()
        _lib.lib.nf_prof_enable(0)
        assert prof["tc_wgrad"][2] == 0 and prof["col_reduce"][2] == 0, prof
        assert all(p.grad is None for p in model.parameters())
        sd64 = fo.cast_state_dict(sd, np.float64)
        z64, _, _ = fo.flow_forward(sd64, 3, x.astype(np.float64), y.astype(np.float64))
        ref = fo.flow_backward(sd64, 3, x.astype(np.float64), y.astype(np.float64), -z64, np.ones(20))
        assert rel_err(grad.cpu().numpy(), ref["dx"]) <= 2e-5
        # ... and a normal training backward on the same model still produces every parameter gradient
        z, _, ld = model(t(x), t(y))
        (-(-0.5 * (z * z).sum(1) + ld).mean()).backward()
        assert all(p.grad is not None for p in model.parameters() if p.requires_grad)


def test_outputs_are_independent_tensors():
    """torch_fcnf.py:134-141 returns fresh tensors: outputs kept across iterations must not be overwritten by the
    next forward (the forward->backward stash they used to alias is pooled and reused)."""
    cfg = dict(D=8, C=64, R=32, n=3, B=64)
    sd = synth.make_state_dict(8, 64, 32, 3, seed=2)
    model = build_model(cfg, sd, "fp32")
    x1, y1 = synth.make_inputs(64, 8, 32, seed=1)
    x2, y2 = synth.make_inputs(64, 8, 32, seed=2)
    _, outs1, _ = model(t(x1), t(y1))          # the z / log_dets graph is dropped: the stash slot is free again
    keep = [o.clone() for o in outs1]
    _, outs2, _ = model(t(x2), t(y2))
    torch.cuda.synchronize()
    for a, b in zip(outs1, keep):
        assert torch.equal(a, b)
    assert not torch.equal(outs1[0], outs2[0])


def test_reference_constructor_call_runs_on_the_tensor_cores():
    """RealNVP(in_channels, channels, cond_channels, n_layers, prior) exactly as torch_nfbc_ant.py:147 writes it
    (no keyword extras) must launch the tcgen05 GEMM kernels -- checked with the library's per-kernel-class profiler."""
    D, C, R, n, B = 8, 256, 64, 2, 512
    prior = torch.distributions.MultivariateNormal(torch.zeros(D, device=dev()), torch.eye(D, device=dev()))
    model = nf_b200.RealNVP(D, C, R, n, prior).to(dev())
    x, y = synth.make_inputs(B, D, R, seed=5)
    _lib.lib.nf_prof_enable(1)
    _lib.prof_collect()
    z, outs, ld = model(t(x), t(y).requires_grad_(True))
    (-(model.prior.log_prob(z) + ld).mean()).backward()
    torch.cuda.synchronize()
    prof = _lib.prof_collect()
    _lib.lib.nf_prof_enable(0)
    assert prof["tc_gemm"][2] >= 2 * n and prof["tc_wgrad"][2] >= n, prof


def _torch_checker(cfg, sd, H1=None, eps=1e-5, dtype=torch.float64):
    from oracle.torch_flow import TorchFlow
    m = TorchFlow(cfg["D"], cfg["C"], cfg["R"], cfg["n"], first_hidden=H1, ln_eps=eps)
    m.load_state_dict({k: torch.tensor(np.asarray(v)) for k, v in sd.items()})
    return m.to(dev()).to(dtype)


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("D,C,R,n,H1,eps,B", [(8, 128, 32, 3, None, 1e-5, 300), (9, 64, 24, 3, None, 1e-5, 50),
                                              (2, 32, 8, 2, 40, 1e-6, 257), (40, 64, 16, 2, None, 1e-5, 64)])
def test_reverse_is_differentiable(D, C, R, n, H1, eps, B, fast):
    """the sampling direction with its log-det, differentiated w.r.t. its input, y and every
    parameter (both directions are differentiated by offline-rl/agents/vinf.py:50,68; flax nf.py:131-148) -- against
    autograd through the float64 checker (oracle/torch_flow.py)."""
    sd = synth.make_state_dict(D, C, R, n, seed=21, H1=H1)
    cfg = dict(D=D, C=C, R=R, n=n, B=B)
    model = build_model(cfg, sd, "fp32", first_hidden=H1, ln_eps=eps, fast_param_grads=fast)
    ref = _torch_checker(cfg, sd, H1, eps)
    # a latent whose reverse lands on kink-free activations: z = forward(x) of safe inputs
    x0, y, _ = synth.make_safe_inputs(sd, n, B, D, R, seed=300 + B, eps=eps)
    with torch.no_grad():
        z64, _, _ = ref(t(x0).double(), t(y).double())
    cx, cl = synth.make_cotangents(B, D, seed=B + 7)
    zt, yt = z64.float().clone().requires_grad_(True), t(y).requires_grad_(True)
    xr, ld = model.reverse(zt, yt, return_logdet=True)
    assert xr.requires_grad and ld.requires_grad
    ((xr * t(cx)).sum() + (ld * t(cl)).sum()).backward()
    zr, yr = z64.float().double().clone().requires_grad_(True), t(y).double().requires_grad_(True)
    xr64, ld64 = ref.reverse(zr, yr, return_logdet=True)
    ((xr64 * t(cx).double()).sum() + (ld64 * t(cl).double()).sum()).backward()
    tol = 1e-4   # the reference's own float32 round trip is 3e-5 .. 1.7e-2
    assert rel_err(xr.detach().cpu().numpy(), xr64.detach().cpu().numpy()) <= tol
    assert rel_err(ld.detach().cpu().numpy(), ld64.detach().cpu().numpy()) <= tol
    assert rel_err(zt.grad.cpu().numpy(), zr.grad.cpu().numpy()) <= tol
    assert rel_err(yt.grad.cpu().numpy(), yr.grad.cpu().numpy()) <= tol
    refg = dict(ref.named_parameters())
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, k
            assert rel_err(p.grad.cpu().numpy(), refg[k].grad.cpu().numpy()) <= 2e-4, k
    # without the log-det, and under no_grad, nothing is recorded
    for p in model.parameters():
        p.grad = None
    x2 = model.reverse(zt.detach().requires_grad_(True), yt.detach())
    assert x2.requires_grad and torch.allclose(x2, xr.detach())
    with torch.no_grad():
        assert not model.reverse(zt, yt).requires_grad


@pytest.mark.parametrize("D,C,R,n,H1,B", [(2, 256, 64, 2, 320, 1000), (8, 128, 32, 3, None, 300), (40, 64, 16, 2, None, 64)])
def test_zero_conditioning_fast_path(D, C, R, n, H1, B):
    """y=None (train_auto_NF_rl.py:328-339 evaluates the flow with cond == 0) computes what
    y = zeros computes -- forward, log-det, gradients, reverse -- without the y part of the first Linear."""
    sd = synth.make_state_dict(D, C, R, n, seed=31, H1=H1)
    cfg = dict(D=D, C=C, R=R, n=n, B=B)
    x, _ = synth.make_inputs(B, D, R, seed=5)
    y0 = torch.zeros(B, R, device=dev())
    dz, dld = synth.make_cotangents(B, D, seed=6)
    res = []
    for yy in (y0, None):
        model = build_model(cfg, sd, "fp32", first_hidden=H1)
        xt = t(x).requires_grad_(True)
        z, outs, ld = model(xt, yy)
        ((z * t(dz)).sum() + (ld * t(dld)).sum()).backward()
        with torch.no_grad():
            xr, rld = model.reverse(z.detach(), yy, return_logdet=True)
            lp = model.log_prob(t(x), yy)
        g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad])
        res.append([v.detach().cpu().numpy() for v in (z, ld, xt.grad, g, xr, rld, lp)])
        if yy is None:   # the y columns of the first Linear's weight gradient are exactly zero
            dc = (D + 1) // 2
            assert float(model.blocks[0].s[0].weight.grad[:, dc:].abs().max()) == 0.0
    for a, b, nm in zip(res[0], res[1], ("z", "ld", "dx", "dparams", "rev", "rev_ld", "log_prob")):
        assert rel_err(b, a) <= 5e-6, (nm, rel_err(b, a))
    sd64 = fo.cast_state_dict(sd, np.float64)
    ref = fo.flow_forward(sd64, n, x.astype(np.float64), np.zeros((B, R)), eps=1e-5)
    assert rel_err(res[1][0], ref[0]) <= 2e-5


@pytest.mark.parametrize("name", ["synth_C2a", "synth_C1a", "synth_C1b", "synth_C3t"])
def test_reduced_precision_mode_contract(name):
    """The optional reduced-precision mode (precision="tf32": one TF32 tensor-core pass on the raw fp32 operands):
    z and log_prob within 2e-2 relative of the float64 reference on the configs inside the stated envelope
    (D <= 64, n_layers <= 12; BASELINE.json north_star), gradients and the inverse stay usable."""
    path = [p for p in CASES if os.path.basename(p) == name + ".npz"][0]
    cfg, sd, x, y, dz, dld, g = load_case(path)
    assert cfg["D"] <= 64 and cfg["n"] <= 12
    model = build_model(cfg, sd, "tf32")
    xt, yt = t(x).requires_grad_(True), t(y).requires_grad_(True)
    _lib.lib.nf_prof_enable(1)
    _lib.prof_collect()
    z, outputs, ld = model(xt, yt)
    (z * t(dz)).sum().add((ld * t(dld)).sum()).backward()
    torch.cuda.synchronize()
    prof = _lib.prof_collect()
    _lib.lib.nf_prof_enable(0)
    assert prof["tc_gemm"][2] > 0 and prof["tc_wgrad"][2] > 0     # the tensor-core kernels ran
    errs = {"z": rel_err(z.detach().cpu().numpy(), g["f64/z"]),
            "log_dets": rel_err(ld.detach().cpu().numpy(), g["f64/log_dets"]),
            "dx": rel_err(xt.grad.cpu().numpy(), g["f64/dx"]), "dy": rel_err(yt.grad.cpu().numpy(), g["f64/dy"])}
    with torch.no_grad():
        nll = -model.log_prob(t(x), t(y)).mean().item()
        rev = model.reverse(t(g["f32/z"]), t(y))
    errs["nll"] = abs(nll - float(g["f64/nll"])) / max(1.0, abs(float(g["f64/nll"])))
    errs["rev"] = rel_err(rev.cpu().numpy(), g["f64/rev"])
    print(f"\n[tf32 mode] {name}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert errs["z"] <= 2e-2 and errs["log_dets"] <= 2e-2 and errs["nll"] <= 2e-2, errs
    # (gradients and the inverse amplify the conditioner error through exp(s) over the blocks: looser bounds, wider
    # still for the 64-dimensional flow whose reference float32 inverse is itself only good to ~1e-3)
    assert errs["dx"] <= 1e-1 and errs["dy"] <= 1e-1 and errs["rev"] <= (1e-1 if cfg["D"] <= 9 else 5e-1), errs
    assert errs["z"] > 1e-6     # ... and it really is the reduced mode


def test_noise_augmentation_kernel_matches_the_philox_oracle():
    """x + 0.1 * randn (torch_nfbc_ant.py:264-265) with the library's Philox4x32-10 kernel: the normals are the ones
    the known-answer-pinned restatement (oracle/philox.py) produces for (seed, offset), odd sizes included; the
    module-level call advances the counter between calls."""
    from oracle import philox
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for n, seed, offset in [(4096 * 8, 3, 0), (1001, 2 ** 40 + 5, 123456789), (3, 9, 2 ** 33)]:
        x = torch.linspace(-1, 1, n, device=dev())
        out = torch.empty_like(x)
        _lib.check(_lib.lib.nf_noise_augment(x.data_ptr(), n, 0.1, seed, offset, out.data_ptr(), st), "nf_noise_augment")
        z = ((out.double() - x.double()) / 0.1).cpu().numpy()
        ref = philox.normals(n, seed, offset)
        assert np.max(np.abs(z - ref)) <= 2e-4, (n, np.max(np.abs(z - ref)))    # fast log / sincos + the fp32 add
    model = build_model(dict(D=8, C=32, R=16, n=2, B=0), synth.make_state_dict(8, 32, 16, 2, seed=1), "fp32")
    x = torch.zeros(512, 8, device=dev())
    a, b = model.augment(x, 0.1, seed=11), model.augment(x, 0.1, seed=11)
    assert not torch.equal(a, b)                                    # fresh noise per call ...
    ref = philox.normals(2 * 4096, 11)
    assert np.max(np.abs(torch.cat([a, b]).reshape(-1).double().cpu().numpy() / 0.1 - ref)) <= 2e-4   # ... one stream
    assert abs(float(a.std()) / 0.1 - 1.0) < 0.05
    # the fused-loss entry point takes the augmentation as an argument
    y = torch.zeros(512, 16, device=dev())
    l0, l1 = model.nll(x, y), model.nll(x, y, noise_std=0.1, seed=11)
    assert torch.isfinite(l1) and float(l0) != float(l1)


@pytest.mark.parametrize("D,C,R,n,B", [(2, 256, 64, 2, 300), (8, 64, 16, 3, 50), (2, 192, 64, 2, 1000)])
def test_flax_semantics_flags(D, C, R, n, B):
    """The flax twin (unsupservised-gcrl/nf.py): first hidden width C + R, eps 1e-6, LayerNorm variance as
    max(0, E[x^2] - E[x]^2), forward -> (z, log_dets), reverse -> (x, log_det).  Parity unpinned upstream (no jax in the
    image): checked against the numpy oracle with the matching variance formula."""
    H1, eps = C + R, 1e-6
    sd = synth.make_state_dict(D, C, R, n, seed=51, H1=H1)
    x, y, _ = synth.make_safe_inputs(sd, n, B, D, R, seed=52, eps=eps)
    dz, dld = synth.make_cotangents(B, D, seed=53)
    m = nf_b200.RealNVP(D, C, R, n, None, first_hidden=H1, ln_eps=eps, ln_fast_variance=True, return_outputs=False)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    m = m.to(dev())
    xt, yt = t(x).requires_grad_(True), t(y).requires_grad_(True)
    res = m(xt, yt)
    assert isinstance(res, tuple) and len(res) == 2            # nf.py:174-179
    z, ld = res
    (z * t(dz)).sum().add((ld * t(dld)).sum()).backward()
    sd64 = fo.cast_state_dict(sd, np.float64)
    fo.LN_FAST_VARIANCE = True
    try:
        ref = fo.flow_backward(sd64, n, x.astype(np.float64), y.astype(np.float64), dz.astype(np.float64),
                               dld.astype(np.float64), eps=eps)
        rx, rld = fo.flow_reverse(sd64, n, ref["z"], y.astype(np.float64), eps=eps, return_logdet=True)
    finally:
        fo.LN_FAST_VARIANCE = False
    assert rel_err(z.detach().cpu().numpy(), ref["z"]) <= 2e-5
    assert rel_err(ld.detach().cpu().numpy(), ref["log_dets"]) <= 2e-5
    assert rel_err(xt.grad.cpu().numpy(), ref["dx"]) <= 2e-5 and rel_err(yt.grad.cpu().numpy(), ref["dy"]) <= 2e-5
    for k, p in m.named_parameters():
        if p.requires_grad:
            assert rel_err(p.grad.cpu().numpy(), ref["grads"][k]) <= 5e-5, k
    with torch.no_grad():
        xr, rl = m.reverse(z.detach(), t(y), return_logdet=True)   # nf.py:131-148: reverse returns its log-det
        z2, ld2 = m(t(x), t(y))
    assert rel_err(xr.cpu().numpy(), rx) <= 1e-4 and rel_err(rl.cpu().numpy(), rld) <= 1e-4
    assert torch.allclose(z2, z.detach(), atol=1e-6) and torch.allclose(ld2, ld.detach(), atol=1e-5)


def test_group_argmin_of_log_prob_matches_the_oracle():
    """The exploration-goal selection of unsupservised-gcrl/train_auto_NF_rl.py:328-333 -- for every group of N candidate
    goals the one with the lowest log N(z) + logdet under masked (all-zero) conditioning -- as one forward + one reduction
    kernel: indices and rows equal argmin over the float64 oracle's log-probabilities (flax-variant network: H1 = C + R,
    eps 1e-6; D = 2 like the goal xy of the reference)."""
    D, C, R, n, G, N = 2, 64, 16, 3, 37, 50
    sd = synth.make_state_dict(D, C, R, n, seed=71, H1=C + R)
    rs = np.random.RandomState(72)
    x = rs.uniform(-1, 1, size=(G, N, D)).astype(np.float32)
    model = nf_b200.RealNVP(D, C, R, n, None, first_hidden=C + R, ln_eps=1e-6)
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    model = model.to(dev())
    sd64 = fo.cast_state_dict(sd, np.float64)
    z, _, ld = fo.flow_forward(sd64, n, x.reshape(G * N, D).astype(np.float64), np.zeros((G * N, R)), eps=1e-6)[:3]
    lp = (-0.5 * (z * z).sum(1) - 0.5 * D * np.log(2 * np.pi) + ld).reshape(G, N)
    want = lp.argmin(1)
    for yy in (None, torch.zeros(G, N, R, device=dev())):
        got, idx = model.argmin_log_prob(t(x), yy, return_index=True)
        idx = idx.cpu().numpy()
        # (a float32 near-tie may pick the runner-up: accept it when the two log-probabilities agree to 1e-5)
        for g in range(G):
            assert idx[g] == want[g] or abs(lp[g, idx[g]] - lp[g, want[g]]) <= 1e-5 * max(1.0, abs(lp[g, want[g]])), (g, idx[g], want[g])
        assert np.array_equal(got.cpu().numpy(), x[np.arange(G), idx])
    # ties go to the smaller index (jnp.argmin): identical candidates
    xt = t(np.repeat(x[:, :1], 4, axis=1).copy())
    _, idx = model.argmin_log_prob(xt, None, return_index=True)
    assert int(idx.abs().max()) == 0


@pytest.mark.parametrize("D,B", [(2, 1000), (8, 4096), (9, 257), (32, 300), (400, 50)])
def test_standard_normal_prior_matches_multivariate_normal(D, B):
    """``nf_b200.StandardNormal(D).log_prob`` (one kernel forward, one backward) against the prior the reference's scripts
    build, ``MultivariateNormal(zeros(D), eye(D))`` (torch_nfbc_ant.py:146), values and gradient, in the loss line's form."""
    rs = np.random.RandomState(D)
    z = torch.tensor(rs.randn(B, D).astype(np.float32), device=dev())
    w = torch.tensor(rs.randn(B).astype(np.float32), device=dev())
    ref_prior = torch.distributions.MultivariateNormal(torch.zeros(D, dtype=torch.float64, device=dev()),
                                                       torch.eye(D, dtype=torch.float64, device=dev()))
    z64 = z.double().requires_grad_(True)
    lp64 = ref_prior.log_prob(z64)
    (lp64 * w.double()).sum().backward()
    zz = z.clone().requires_grad_(True)
    prior = nf_b200.StandardNormal(D, dev())
    lp = prior.log_prob(zz)
    (lp * w).sum().backward()
    assert lp.shape == (B,)
    assert rel_err(lp.detach().cpu().numpy(), lp64.detach().cpu().numpy()) <= 2e-6
    assert rel_err(zz.grad.cpu().numpy(), z64.grad.cpu().numpy()) <= 2e-6
    assert prior.sample((7,)).shape == (7, D) and prior.sample((7,)).device.type == "cuda"
    # leading batch dimensions and a view with an offset (rows of a larger tensor)
    lp3 = prior.log_prob(z[1:].reshape(1, B - 1, D))
    assert lp3.shape == (1, B - 1) and torch.allclose(lp3[0], lp[1:].detach(), rtol=1e-6, atol=1e-6)
