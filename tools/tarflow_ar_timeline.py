"""Cycle breakdown of the persistent autoregressive-inverse kernel (csrc/tarflow_ar.cu): per CTA, cycles spent in the
embedding, the five phases of a layer, the tail and the grid barriers over one block's loop.  python tools/tarflow_ar_timeline.py [B]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from nf_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
m = nf_b200.TransformerRealNVP(8, 256, 64, 64, 4, 4, None).to(dev)
z = torch.randn(B, 8, 1, device=dev); y = torch.randn(B, 1, 64, device=dev)
_lib.lib.nf_graphs_enable(0)
m.reverse(z, y); torch.cuda.synchronize()
buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
_lib.lib.nf_debug_tc_timeline.argtypes = [ctypes.c_void_p]
_lib.lib.nf_debug_tc_timeline(buf.data_ptr())
m.reverse(z, y); torch.cuda.synchronize()
_lib.lib.nf_debug_tc_timeline(None)
t = buf.view(148, 8)[:128].double().cpu()       # (last block's launch overwrote the earlier ones)
names = ["embed", "P1 LN1+qkv", "P2 attention", "P3 proj", "P4 LN2+fc1+gelu", "P5 fc2", "tail", "grid barriers (incl. waiting)"]
tot = t.sum(1).mean()
print(f"B={B}: one block = {tot:.0f} cycles = {tot / 1.9e3:.1f} us (mean over 128 CTAs)")
for i, n in enumerate(names):
    print(f"  {n:32s} mean {t[:, i].mean():9.0f}  min {t[:, i].min():9.0f}  max {t[:, i].max():9.0f}  ({100 * t[:, i].mean() / tot:4.1f} %)")
