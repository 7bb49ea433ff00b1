"""In-kernel timeline (nf_debug_tc_timeline, 8 clock64 stamps per CTA) of the LAST tcgen05 GEMM launch of a C2a training
step: the batched dW1 weight-gradient GEMM (grid (2, 1, 72); its CTAs are rows 0..143 of the buffer).
    NF_B200_DEFER_DY=0 python tools/wgrad_timeline.py      (keeps the dy GEMMs inside the block loop, before that launch)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from nf_b200 import _lib
from oracle import synth
lib = _lib.lib
lib.nf_debug_tc_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
lib.nf_graphs_enable(0)
cfg = dict(synth.CONFIGS["C2a"]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps)
m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
x, y = synth.make_inputs(4096, cfg["D"], cfg["R"])
x = torch.tensor(x, device=dev); y = torch.tensor(y, device=dev)
plist = [p for p in m.parameters() if p.requires_grad]

def step():
    for p in plist: p.grad = None
    (-m.log_prob(x, y.detach().requires_grad_(True)).mean()).backward()

for _ in range(3): step()
torch.cuda.synchronize()
NAMES = ["prologue", "first tile", "mainloop(split)", "mma drain", "tmem->regs (1st chunk)", "1st chunk stores", "rest of epilogue"]
for rows, label in ((144, "dW1 GEMM  k_tc_gemm<64,12,MN> grid (2,1,72): rows 0..143 of the last launch"),):
    buf = torch.zeros(4096 * 8, dtype=torch.int64, device=dev)
    lib.nf_debug_tc_timeline(buf.data_ptr())
    step()
    torch.cuda.synchronize()
    lib.nf_debug_tc_timeline(None)
    t = buf.view(-1, 8).cpu()[:rows]
    t = t[t[:, 0] != 0]
    print(label, "-", len(t), "CTAs; mean cycles per phase:")
    d = (t[:, 1:8] - t[:, :7]).double()
    for i, n in enumerate(NAMES):
        if (t[:, i + 1] == 0).all() or (t[:, i] == 0).all(): continue
        v = d[:, i]
        print(f"    {n:28s} {v.mean():9.0f}  (min {v.min():7.0f} max {v.max():7.0f})")
    tot = (t[:, 7] - t[:, 0]).double()
    print(f"    total in CTA {tot.mean():9.0f} cycles = {tot.mean() / 1.965e3:.2f} us; first start -> last end "
          f"{(t[:, 7].max() - t[:, 0].min()) / 1.965e3:.2f} us")
