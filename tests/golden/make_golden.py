"""Generate the golden fixtures in this directory from the REAL reference module.

Run in the build container only (it imports /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

Two families of known-answer tests (the reference ships none):

* ``refinit_*``: the reference's own constructor/initialiser (orthogonal init + LU,
  torch_fcnf.py:15-31, default nn.Linear/LayerNorm init) under torch.manual_seed(0), with the
  zero-initialised last layers perturbed.  Small shapes; the whole
  state_dict is stored.
* ``synth_*``: BASELINE.json shapes.  Weights/inputs come from oracle/synth.py (bit-reproducible
  from a seed, so they are NOT stored), are loaded into the reference module with
  load_state_dict, and the reference's outputs are stored: z, outputs, log_dets, reverse(z),
  dx, dy and per-parameter gradient probes.  Every quantity is stored twice: from the
  module in float32 (what the reference computes) and in float64 (the arbiter, BASELINE.md
  section 3).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

REF = "/root/reference/il-and-cil/utils/torch_fcnf.py"
spec = importlib.util.spec_from_file_location("ref_torch_fcnf", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

N_PROBE = 32


def probe_indices(numel, key):
    rs = np.random.RandomState(abs(hash_str(key)) % (2**31))
    return rs.randint(0, numel, size=N_PROBE)


def hash_str(s):
    h = 0
    for ch in s:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def run_reference(model, x, y, dz, dld, dtype):
    """forward, reverse and backward of the unmodified reference module."""
    model = model.to(dtype)
    xt = torch.tensor(x, dtype=dtype, requires_grad=True)
    yt = torch.tensor(y, dtype=dtype, requires_grad=True)
    for p in model.parameters():
        p.grad = None
    z, outputs, log_dets = model(xt, yt)
    obj = (z * torch.tensor(dz, dtype=dtype)).sum() + (log_dets * torch.tensor(dld, dtype=dtype)).sum()
    obj.backward()
    D = x.shape[1]
    nll = -((-0.5 * (z * z).sum(1) - 0.5 * D * np.log(2 * np.pi)) + log_dets).mean()
    with torch.no_grad():
        seq = model.reverse(z.detach(), yt.detach(), return_sequence=True)
    res = dict(
        z=z.detach().numpy(), log_dets=log_dets.detach().numpy(),
        outputs=np.stack([o.detach().numpy() for o in outputs]),
        rev=seq[-1].numpy(), rev_seq=np.stack([s.numpy() for s in seq]),
        dx=xt.grad.numpy(), dy=yt.grad.numpy(), nll=np.asarray(nll.item()),
    )
    grads = {k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    return res, grads


def build(D, C, R, n):
    prior = torch.distributions.MultivariateNormal(torch.zeros(D), torch.eye(D))
    return ref.RealNVP(D, C, R, n, prior)


def make_refinit(name, D, C, R, n, B):
    torch.manual_seed(0)
    model = build(D, C, R, n)
    with torch.no_grad():
        for blk in model.blocks:
            for net in (blk.s, blk.t):
                net[6].weight.normal_(0.0, 0.2 / np.sqrt(C))
                net[6].bias.normal_(0.0, 0.2 / np.sqrt(C))
                # exercise the LayerNorm affine parameters as well
                net[2].weight.add_(0.1 * torch.randn_like(net[2].weight))
                net[2].bias.add_(0.1 * torch.randn_like(net[2].bias))
                net[5].weight.add_(0.1 * torch.randn_like(net[5].weight))
                net[5].bias.add_(0.1 * torch.randn_like(net[5].bias))
    sd32 = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    x, y, xseed = synth.make_safe_inputs(sd32, n, B, D, R, seed=1234)
    dz, dld = synth.make_cotangents(B, D, seed=4321)
    out = dict(cfg=np.asarray([D, C, R, n, B, 0, xseed]))
    for k, v in sd32.items():
        out["sd/" + k] = v
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = build(D, C, R, n)
        m.load_state_dict({k: torch.tensor(v) for k, v in sd32.items()})
        res, grads = run_reference(m, x.astype(np.float64 if dt == torch.float64 else np.float32),
                                   y.astype(np.float64 if dt == torch.float64 else np.float32), dz, dld, dt)
        for k, v in res.items():
            out[f"{tag}/{k}"] = v
        for k, g in grads.items():
            out[f"{tag}/grad/{k}"] = g
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok, input seed", xseed)


def make_synth(name, cfgname, B, n_override=None, seed=0):
    cfg = dict(synth.CONFIGS[cfgname])
    cfg.pop("ln_eps")
    H1 = cfg.pop("H1")
    assert H1 == cfg["C"], "the torch reference has H1 == C"
    if n_override:
        cfg["n"] = n_override
    D, C, R, n = cfg["D"], cfg["C"], cfg["R"], cfg["n"]
    sd = synth.make_state_dict(D, C, R, n, seed=seed)
    x, y, xseed = synth.make_safe_inputs(sd, n, B, D, R, seed=1234)
    dz, dld = synth.make_cotangents(B, D, seed=4321)
    out = dict(cfg=np.asarray([D, C, R, n, B, seed, xseed]))
    for tag, dt, npdt in (("f32", torch.float32, np.float32), ("f64", torch.float64, np.float64)):
        m = build(D, C, R, n)
        m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
        res, grads = run_reference(m, x.astype(npdt), y.astype(npdt), dz, dld, dt)
        for k, v in res.items():
            if k == "rev_seq":
                continue
            out[f"{tag}/{k}"] = v
        for k, g in grads.items():
            idx = probe_indices(g.size, k)
            out[f"{tag}/gprobe/{k}"] = np.concatenate(
                [[g.sum(dtype=np.float64), (g.astype(np.float64) ** 2).sum()], g.reshape(-1)[idx].astype(np.float64)])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok, input seed", xseed)


if __name__ == "__main__":
    torch.set_num_threads(8)
    make_refinit("refinit_d9", D=9, C=32, R=16, n=3, B=37)
    make_refinit("refinit_d2", D=2, C=32, R=8, n=2, B=20)
    make_refinit("refinit_d8", D=8, C=64, R=32, n=2, B=50)
    make_synth("synth_C1a", "C1a", B=64)
    make_synth("synth_C1b", "C1b", B=48)
    make_synth("synth_C1c", "C1c", B=40)
    make_synth("synth_C2a", "C2a", B=257)
    make_synth("synth_C3t", "C3t", B=130)
    make_synth("synth_C4", "C4", B=33)
    make_synth("synth_C5", "C5", B=24)
