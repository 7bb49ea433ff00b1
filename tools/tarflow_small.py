"""Autoregressive sampling of the transformer flow at evaluation-sized batches, tensor-core vs FFMA GEMMs.
python tools/tarflow_small.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from nf_b200 import _lib
dev = torch.device("cuda:0")
for prec in ("fp32", "fp32_simt"):
    m = nf_b200.TransformerRealNVP(8, 256, 64, 64, 4, 4, None, precision=prec).to(dev)
    for B in (20, 256, 1024):
        z = torch.randn(B, 8, 1, device=dev); y = torch.randn(B, 1, 64, device=dev)
        for _ in range(3): m.reverse(z, y)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m.reverse(z, y)
        e1.record(); torch.cuda.synchronize()
        print(f"{prec} B={B}: reverse {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
