"""Fixture loading for the conditioning-encoder tests (row f2)."""
from __future__ import annotations

import glob
import os

import numpy as np

from oracle import synth
from tests.helpers import GOLDEN_DIR

ENC_CASES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "encoder", "enc_*.npz")))


def load_encoder_case(path):
    """Returns (cfg, state_dict, g, dout, golden npz)."""
    gold = np.load(path)
    input_size, rep_size, B, seed, xseed = [int(v) for v in gold["cfg"]]
    sd = synth.make_encoder_state_dict(input_size, rep_size, seed=seed)
    g, dout = synth.make_encoder_inputs(B, input_size, rep_size, seed=xseed)
    return dict(input_size=input_size, rep_size=rep_size, B=B), sd, g, dout, gold
