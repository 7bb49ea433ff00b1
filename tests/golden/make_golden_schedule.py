"""Golden learning-rate sequences from the REAL reference CosineLRSchedule (il-and-cil/utils/torch_utils.py:21-49).
Build container only:  python tests/golden/make_golden_schedule.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_torch_utils", "/root/reference/il-and-cil/utils/torch_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [(5, 40, 1e-6, 3e-4), (1, 8, 1e-4, 1e-3), (10, 12, 0.0, 1.0)]
out = {"cases": np.array(CASES, dtype=np.float64)}
for i, (warm, total, lo, hi) in enumerate(CASES):
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.5)
    sch = ref.CosineLRSchedule(opt, int(warm), int(total), lo, hi)
    ret, lrs = [], [opt.param_groups[0]["lr"]]
    for _ in range(int(total) + 3):          # three steps past the end: the reference stops updating the optimizer there
        ret.append(sch.step())
        lrs.append(opt.param_groups[0]["lr"])
    out[f"ret{i}"], out[f"lr{i}"] = np.array(ret), np.array(lrs)
np.savez(os.path.join(HERE, "schedule", "cosine.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
