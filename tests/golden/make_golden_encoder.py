"""Golden fixtures for the conditioning encoder, produced by the REAL reference
module GEncoder (/root/reference/il-and-cil/utils/torch_fcnf.py:154-190).  Build container only:

    python tests/golden/make_golden_encoder.py

Weights and inputs come from oracle/synth.py (reproducible from a seed, not stored); stored are the reference's
output, input gradient and per-parameter gradient probes, from the module in float32 and in float64 (arbiter).
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
from tests.helpers import grad_probe  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_torch_fcnf", "/root/reference/il-and-cil/utils/torch_fcnf.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

BATCH = {"gcbc": 33, "kitchen": 20, "ant": 16}


def run(model, g, dout, dtype):
    model = model.to(dtype)
    for p in model.parameters():
        p.grad = None
    gt = torch.tensor(g, dtype=dtype, requires_grad=True)
    out = model(gt)
    (out * torch.tensor(dout, dtype=dtype)).sum().backward()
    res = {"out": out.detach().numpy(), "dg": gt.grad.numpy()}
    for k, p in model.named_parameters():
        res["gprobe/" + k] = grad_probe(p.grad.numpy(), k)
    return res


def main():
    torch.manual_seed(0)
    for name, cfg in synth.ENCODER_CONFIGS.items():
        B = BATCH[name]
        sd = synth.make_encoder_state_dict(cfg["input_size"], cfg["rep_size"], seed=0)
        g, dout = synth.make_encoder_inputs(B, cfg["input_size"], cfg["rep_size"])
        out = {"cfg": np.array([cfg["input_size"], cfg["rep_size"], B, 0, 77])}
        for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
            model = ref.GEncoder(input_size=cfg["input_size"], rep_size=cfg["rep_size"])
            model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
            for k, v in run(model, g, dout, dtype).items():
                out[f"{tag}/{k}"] = v
        path = os.path.join(HERE, "encoder", f"enc_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
