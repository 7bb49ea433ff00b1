"""Shared helpers for the parity tests."""
from __future__ import annotations

import glob
import os

import numpy as np

from oracle import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_PROBE = 32


def golden_files(prefix=""):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def rel_err(a, ref):
    """max|a-ref| / max|ref| -- the error measure of BASELINE.md section 3."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(float(np.max(np.abs(ref))), 1e-30)
    return float(np.max(np.abs(a - ref))) / scale


def tol_vs_arbiter(ref32, ref64, floor=1e-5):
    """fp32-parity tolerance: err(ours, ref64) <= max(1e-5, 2 err(ref32, ref64))."""
    return max(floor, 2.0 * rel_err(ref32, ref64))


def hash_str(s):
    h = 0
    for ch in s:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def probe_indices(numel, key):
    rs = np.random.RandomState(abs(hash_str(key)) % (2**31))
    return rs.randint(0, numel, size=N_PROBE)


def grad_probe(g, key):
    g = np.asarray(g)
    idx = probe_indices(g.size, key)
    return np.concatenate([[g.sum(dtype=np.float64), (g.astype(np.float64) ** 2).sum()],
                           g.reshape(-1)[idx].astype(np.float64)])


def load_case(path):
    """Returns (cfg dict, sd, x, y, dz, dld, golden npz)."""
    g = np.load(path)
    name = os.path.basename(path)
    D, C, R, n, B, seed, xseed = [int(v) for v in g["cfg"]]
    if name.startswith("refinit"):
        sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    else:
        sd = synth.make_state_dict(D, C, R, n, seed=seed)
    x, y = synth.make_inputs(B, D, R, seed=xseed)   # kink-free seed chosen by make_golden.py
    dz, dld = synth.make_cotangents(B, D, seed=4321)
    return dict(D=D, C=C, R=R, n=n, B=B), sd, x, y, dz, dld, g


def probe_err(ours_probe, ref_probe, numel):
    """Error of a gradient probe (sum, sumsq, 32 samples) normalised by the gradient's rms
    (a 32-entry sample's own max can be far below the tensor's scale)."""
    rms = max(np.sqrt(ref_probe[1] / numel), 1e-30)
    scale = max(float(np.max(np.abs(ref_probe[2:]))), rms)
    e_samples = float(np.max(np.abs(ours_probe[2:] - ref_probe[2:]))) / scale
    e_sum = abs(ours_probe[0] - ref_probe[0]) / (rms * np.sqrt(numel))
    e_sq = abs(ours_probe[1] - ref_probe[1]) / max(ref_probe[1], 1e-30)
    return max(e_samples, e_sum, e_sq)
