"""nf_b200.CosineLRSchedule against learning-rate sequences produced by the real reference class
(tests/golden/make_golden_schedule.py; il-and-cil/utils/torch_utils.py:21-49)."""
import os

import numpy as np
import torch

import nf_b200
from tests.helpers import GOLDEN_DIR


def test_cosine_schedule_matches_reference_sequences():
    g = np.load(os.path.join(GOLDEN_DIR, "schedule", "cosine.npz"))
    for i, (warm, total, lo, hi) in enumerate(g["cases"]):
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.5)
        sch = nf_b200.CosineLRSchedule(opt, int(warm), int(total), float(lo), float(hi))
        lrs, ret = [opt.param_groups[0]["lr"]], []
        for _ in range(int(total) + 3):
            ret.append(sch.step())
            lrs.append(opt.param_groups[0]["lr"])
        np.testing.assert_allclose(ret, g[f"ret{i}"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(lrs, g[f"lr{i}"], rtol=1e-12, atol=1e-15)
        assert "counter" in sch.state_dict() and float(sch.counter) == int(total) + 3
